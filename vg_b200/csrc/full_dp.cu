// full_dp.cu — B3 stage seam: Aligner::align(alignment, graph, topological_order), the full
// (unbanded) local alignment of a read against a DAG that vg runs through GSSW
// (aligner.cpp:571-626 -> align_internal :344-564 -> gssw_graph_fill_pinned / gssw_graph_trace_back).
// vg giraffe reaches it as the rescue fallback (fix_dozeu_score, minimizer_mapper.cpp:3502-3517)
// and with --rescue-algorithm gssw (:3390).  The recurrence, end-cell and traceback rules are the
// ones stated in oracle/full_dp.cpp (gssw itself is absent from the reference tree).
//
// One warp per problem.  A DP column is swept in 32-cell chunks along the query (lane = query
// offset), the insertion chain by a max-plus prefix scan over shuffles; the two live columns sit in
// shared memory, every node's last column (for its successors), the predecessor choices of first
// columns and one traceback byte per cell in a per-problem HBM workspace whose size the host knows.
#include <cub/cub.cuh>
#include "giraffe_b200.h"
#include "device_state.cuh"
#include "align.cuh"

#include <algorithm>
#include <vector>

namespace gb {

constexpr int SW_WARPS = 4;

struct SwBatch {
    const uint32_t* node; const uint64_t* node_off;         // problem p: nodes node_off[p] .. node_off[p+1] (topological order)
    const uint32_t* pred; const uint64_t* pred_off;         // CSR over all nodes (global node index): predecessor indices within the problem
    const uint8_t* query; const uint64_t* query_off;
    uint32_t n; uint32_t Lc;
    int32_t* score; gb_mapping* maps; uint32_t* edits; uint32_t* n_maps; uint32_t* n_edits; uint8_t* status;
    uint32_t map_cap, edit_cap;
    // per-problem workspace offsets (host prefix sums)
    int32_t* lastH; int32_t* lastE; uint8_t* argH; uint8_t* argE; const uint64_t* col_off;    // node-granular: (global node index) * W_p via col_off[p] + u * W
    uint8_t* tb; const uint64_t* tb_off;                    // byte offset of problem p's traceback matrix
    uint32_t* col_start; uint32_t* seq_off; uint32_t* seq_len;   // per global node: first traceback column, sequence
    uint32_t* steps; const uint64_t* step_off;
    uint32_t* work_counter;
};

// tb byte: bits 0-1 source of H (0 diagonal, 1 deletion, 2 insertion, 3 none: H == 0),
//          bit 2 deletion opened here, bit 3 insertion opened here, bit 4 the diagonal starts the alignment
__device__ inline void sw_problem(const DevIndex& ix, const DevScores& sc, const SwBatch& b, uint32_t p,
                                  uint8_t* q, DpSmem dps, int32_t& score_out, PathBuf& out, uint32_t& status) {
    const int lane = lane_id();
    const uint64_t n0 = b.node_off[p]; const uint32_t N = (uint32_t)(b.node_off[p + 1] - n0);
    const uint64_t q0 = b.query_off[p]; const uint32_t m = (uint32_t)(b.query_off[p + 1] - q0);
    const uint32_t W = m + 1;
    const int32_t go = sc.gap_open, ge = sc.gap_extend, bonus = sc.full_length_bonus;
    score_out = 0; pb_reset(out);
    if (N == 0 || m == 0) return;
    for (uint32_t i = lane; i < m; i += 32) q[i] = b.query[q0 + i];
    // node table: sequence and first traceback column
    if (lane == 0) {
        uint32_t cols = 0;
        for (uint32_t u = 0; u < N; u++) {
            const gb_node_rec nr = load_node(ix, b.node[n0 + u]);
            b.seq_off[n0 + u] = nr.seq_off; b.seq_len[n0 + u] = nr.len; b.col_start[n0 + u] = cols;
            cols += nr.len;
        }
    }
    __syncwarp();
    int32_t* lastH = b.lastH + b.col_off[p]; int32_t* lastE = b.lastE + b.col_off[p];
    uint8_t* argH = b.argH + b.col_off[p]; uint8_t* argE = b.argE + b.col_off[p];
    uint8_t* tb = b.tb + b.tb_off[p];

    // this lane's best end cell: (candidate score, node, column, query offset, attached right end)
    int32_t lane_best = 0; uint32_t lane_u = 0, lane_c = 0, lane_j = 0; bool lane_end = false;
    for (uint32_t u = 0; u < N; u++) {
        const uint32_t len = b.seq_len[n0 + u], soff = b.seq_off[n0 + u];
        // merged incoming column
        const uint64_t pb0 = b.pred_off[n0 + u]; const uint32_t np = (uint32_t)(b.pred_off[n0 + u + 1] - pb0);
        for (uint32_t j = lane; j < W; j += 32) {
            int32_t h = DP_NEG, e = DP_NEG; uint32_t ah = 0xff, ae = 0xff;
            for (uint32_t pi = 0; pi < np; pi++) {
                const uint32_t pu = b.pred[pb0 + pi];
                const int32_t ph = lastH[(size_t)pu * W + j], pe = lastE[(size_t)pu * W + j];
                if (ph > h) { h = ph; ah = pi; }
                if (pe > e) { e = pe; ae = pi; }
            }
            dps.Hp[j] = h; dps.Ep[j] = e; argH[(size_t)u * W + j] = (uint8_t)ah; argE[(size_t)u * W + j] = (uint8_t)ae;
        }
        __syncwarp();
        for (uint32_t c = 0; c < len; c++) {
            const uint8_t r = __ldg(ix.seq + soff + c);
            const bool r_ok = is_acgt(r);
            uint8_t* tbcol = tb + (size_t)(b.col_start[n0 + u] + c) * W;
            int32_t carry = INT_MIN, prevH_last = DP_NEG, prev_ph_last = DP_NEG;
            for (uint32_t jb = 0; jb < W; jb += 32) {
                const uint32_t j = jb + lane;
                const bool in = j < W;
                int32_t ph = DP_NEG, pe = DP_NEG;
                if (in) { ph = dps.Hp[j]; pe = dps.Ep[j]; }
                int32_t phm1 = __shfl_up_sync(FULL, ph, 1);
                if (lane == 0) phm1 = prev_ph_last;
                prev_ph_last = __shfl_sync(FULL, ph, 31);
                int32_t d = DP_NEG, e = DP_NEG;
                bool fresh = false;
                if (in && j > 0) {
                    const uint8_t qc = q[j - 1];
                    const int32_t s = (!r_ok || !is_acgt(qc)) ? 0 : (qc == r ? sc.match : -sc.mismatch);
                    fresh = j == 1 || !(phm1 > 0);
                    d = (j == 1 ? bonus : max(phm1, 0)) + s;
                    if (ph > 0) e = ph - go;
                    if (pe > DP_NEG) e = max(e, pe - ge);
                    if (e <= 0) e = DP_NEG;
                }
                const int32_t hprime = max(d, e);
                // insertion chain over the cells worth keeping (H' > 0)
                int32_t g = (in && j > 0 && hprime > 0) ? hprime + (int32_t)j * ge : INT_MIN;
                int32_t incl = g;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int32_t t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl = max(incl, t); }
                int32_t excl = __shfl_up_sync(FULL, incl, 1);
                if (lane == 0) excl = INT_MIN;
                excl = max(excl, carry);
                carry = max(carry, __shfl_sync(FULL, incl, 31));
                int32_t f = DP_NEG;
                if (in && j > 0 && excl > INT_MIN) f = excl - go - (int32_t)(j - 1) * ge;
                if (f <= 0) f = DP_NEG;
                int32_t h = 0;
                if (in && j > 0) h = max(max(d, 0), max(e, f));
                int32_t hm1 = __shfl_up_sync(FULL, h, 1);
                if (lane == 0) hm1 = prevH_last;
                prevH_last = __shfl_sync(FULL, h, 31);
                if (in) {
                    uint8_t tbv;
                    if (h <= 0) tbv = 3; else if (d == h) tbv = 0; else if (e == h) tbv = 1; else tbv = 2;
                    if (ph > 0 && e == ph - go) tbv |= 4;
                    if (j > 0 && hm1 > 0 && f == hm1 - go) tbv |= 8;
                    if (fresh) tbv |= 16;
                    dps.Hc[j] = h; dps.Ec[j] = e; tbcol[j] = tbv;
                    if (j > 0) {
                        int32_t cand = h; bool end_diag = false;
                        if (j == m && d + bonus >= h) { cand = d + bonus; end_diag = true; }
                        if (cand > lane_best) { lane_best = cand; lane_u = u; lane_c = c; lane_j = j; lane_end = end_diag; }
                    }
                }
            }
            __syncwarp();
            int32_t* t1p = dps.Hp; dps.Hp = dps.Hc; dps.Hc = t1p;
            int32_t* t2p = dps.Ep; dps.Ep = dps.Ec; dps.Ec = t2p;
        }
        for (uint32_t j = lane; j < W; j += 32) { lastH[(size_t)u * W + j] = dps.Hp[j]; lastE[(size_t)u * W + j] = dps.Ep[j]; }
        __syncwarp();
    }
    // first maximum in (node, column, query offset) order
    int32_t best = lane_best; uint32_t bu = lane_u, bc = lane_c, bj = lane_j; uint32_t bend = lane_end ? 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int32_t ob = __shfl_xor_sync(FULL, best, o);
        const uint32_t ou = __shfl_xor_sync(FULL, bu, o), oc = __shfl_xor_sync(FULL, bc, o), oj = __shfl_xor_sync(FULL, bj, o), oe = __shfl_xor_sync(FULL, bend, o);
        const bool take = ob > best || (ob == best && (ou < bu || (ou == bu && (oc < bc || (oc == bc && oj < bj)))));
        if (take) { best = ob; bu = ou; bc = oc; bj = oj; bend = oe; }
    }
    if (best <= 0) return;
    score_out = best;

    // ---- traceback (uniform): steps end -> start as (node << 8) | op, op 0 match, 1 mismatch, 2 insertion, 3 deletion
    uint32_t* steps = b.steps + b.step_off[p];
    uint32_t n_steps = 0, start_j = 0;
    {
        uint32_t u = bu, c = bc, j = bj;
        int state = bend ? 3 : 0;
        while (true) {
            const uint8_t tbv = tb[(size_t)(b.col_start[n0 + u] + c) * W + j];
            if (state == 0) {
                const uint32_t src = tbv & 3u;
                if (src == 3) { start_j = j; break; }
                state = src == 0 ? 3 : (src == 1 ? 1 : 2);
                continue;
            }
            if (state == 3) {
                const uint8_t qc = q[j - 1], r = __ldg(ix.seq + b.seq_off[n0 + u] + c);
                if (lane == 0) steps[n_steps] = (u << 8) | (qc == r ? 0u : 1u);
                n_steps++;
                j--;
                if (tbv & 16u) { start_j = j; break; }
                if (c > 0) c--; else { u = b.pred[b.pred_off[n0 + u] + argH[(size_t)u * W + j]]; c = b.seq_len[n0 + u] - 1; }
                state = 0;
                continue;
            }
            if (state == 1) {
                if (lane == 0) steps[n_steps] = (u << 8) | 3u;
                n_steps++;
                const bool open = (tbv & 4u) != 0;
                if (c > 0) c--; else { u = b.pred[b.pred_off[n0 + u] + (open ? argH[(size_t)u * W + j] : argE[(size_t)u * W + j])]; c = b.seq_len[n0 + u] - 1; }
                state = open ? 0 : 1;
                continue;
            }
            if (lane == 0) steps[n_steps] = (u << 8) | 2u;
            n_steps++;
            const bool open = (tbv & 8u) != 0;
            j--;
            state = open ? 0 : 2;
        }
    }
    __syncwarp();

    // ---- steps -> mappings; soft clips as insertion edits on the first / last mapping (aligner.cpp:150-241)
    if (lane == 0) {
        uint32_t query_offset = start_j;
        int64_t si = (int64_t)n_steps - 1;
        bool first = true;
        while (si >= 0) {
            const uint32_t nd = steps[si] >> 8;
            // columns this mapping consumes, to place its offset
            uint32_t cols = 0; int64_t k = si;
            while (k >= 0 && (steps[k] >> 8) == nd) { if ((steps[k] & 0xffu) != 2u) cols++; k--; }
            const uint32_t end_col = k < 0 ? bc + 1 : b.seq_len[n0 + nd];
            pb_add_mapping(out, nd, end_col - cols);
            if (first && start_j > 0) pb_add_edit(out, edit_word(GB_EDIT_INS, start_j, 0));
            first = false;
            uint32_t cur = 0xff, run = 0;
            auto flush = [&]() {
                if (cur == 1) { for (uint32_t x = 0; x < run; x++) { pb_add_edit(out, edit_word(GB_EDIT_SUB, 1, base2(q[query_offset]))); query_offset++; } }
                else if (run > 0) {
                    if (cur == 0) { pb_add_edit(out, edit_word(GB_EDIT_MATCH, run, 0)); query_offset += run; }
                    else if (cur == 2) { pb_add_edit(out, edit_word(GB_EDIT_INS, run, 0)); query_offset += run; }
                    else if (cur == 3) { pb_add_edit(out, edit_word(GB_EDIT_DEL, run, 0)); }
                }
            };
            while (si > k) {
                const uint32_t op = steps[si] & 0xffu;
                if (op == cur) run++; else { if (cur != 0xff) flush(); cur = op; run = 1; }
                si--;
            }
            if (cur != 0xff) flush();
        }
        if (out.n_maps > 0 && query_offset != m) pb_add_edit(out, edit_word(GB_EDIT_INS, m - query_offset, 0));
    }
    __syncwarp();
    out.n_maps = __shfl_sync(FULL, out.n_maps, 0); out.n_edits = __shfl_sync(FULL, out.n_edits, 0);
    out.overflow = __shfl_sync(FULL, (int)out.overflow, 0) != 0;
    if (out.overflow) status = GB_ITEM_OUT_FULL;
}

__global__ void __launch_bounds__(SW_WARPS * 32)
sw_kernel(DevIndex ix, DevScores sc, SwBatch b) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t W = b.Lc + 1;
    const size_t per_warp = (((size_t)b.Lc + 15) & ~(size_t)15) + (size_t)W * 4 * 4 + 64;
    uint8_t* base = smem + (size_t)warp * per_warp;
    uint8_t* q = base;
    int32_t* cols = reinterpret_cast<int32_t*>(base + (((size_t)b.Lc + 15) & ~(size_t)15));
    DpSmem dps; dps.Hp = cols; dps.Ep = cols + W; dps.Hc = cols + 2 * W; dps.Ec = cols + 3 * W;
    while (true) {
        uint32_t p = 0;
        if (lane == 0) p = atomicAdd(b.work_counter, 1u);
        p = __shfl_sync(FULL, p, 0);
        if (p >= b.n) break;
        uint32_t status = GB_ITEM_OK; int32_t score = 0;
        PathBuf out; out.maps = b.maps + (size_t)p * b.map_cap; out.edits = b.edits + (size_t)p * b.edit_cap;
        out.map_cap = b.map_cap; out.edit_cap = b.edit_cap; pb_reset(out);
        sw_problem(ix, sc, b, p, q, dps, score, out, status);
        if (lane == 0) { b.score[p] = status == GB_ITEM_OK ? score : 0; b.n_maps[p] = out.n_maps; b.n_edits[p] = out.n_edits; b.status[p] = (uint8_t)status; }
        __syncwarp();
    }
}

} // namespace gb

using namespace gb;

extern "C" int gb_sw_batch(gb_device* d, uint32_t n,
                           const uint32_t* node, const uint64_t* node_off, const uint32_t* pred, const uint64_t* pred_off,
                           const uint8_t* query, const uint64_t* query_off, uint32_t map_cap, uint32_t edit_cap,
                           int32_t* score, gb_mapping* maps, uint32_t* edits, uint32_t* n_maps, uint32_t* n_edits, uint8_t* status) {
    if (!d || !node || !node_off || !pred || !pred_off || !query || !query_off || !score || !maps || !edits || !n_maps || !n_edits ||
        !status || map_cap == 0 || edit_cap == 0) return GB_ERR_ARG;
    if (n == 0) return GB_OK;
    GB_CUDA(cudaSetDevice(d->device));
    const uint64_t total_nodes = node_off[n], total_pred = pred_off[total_nodes], total_q = query_off[n];
    // validate the DAG shape (predecessors precede, known nodes, at most 255 predecessors) and size the workspaces
    uint32_t max_q = 0;
    std::vector<uint64_t> col_off(n + 1, 0), tb_off(n + 1, 0), step_off(n + 1, 0);
    for (uint32_t p = 0; p < n; p++) {
        const uint32_t m = (uint32_t)(query_off[p + 1] - query_off[p]);
        max_q = std::max(max_q, m);
        uint64_t bases = 0;
        const uint64_t N = node_off[p + 1] - node_off[p];
        for (uint64_t u = 0; u < N; u++) {
            const uint32_t v = node[node_off[p] + u];
            if (v < 2 || v >= d->ix.n_nodes) { g_last_error = "gb_sw_batch: unknown node"; return GB_ERR_ARG; }
            const uint64_t g = node_off[p] + u;
            if (pred_off[g + 1] - pred_off[g] > 255) { g_last_error = "gb_sw_batch: more than 255 predecessors"; return GB_ERR_ARG; }
            for (uint64_t e = pred_off[g]; e < pred_off[g + 1]; e++)
                if (pred[e] >= u) { g_last_error = "gb_sw_batch: nodes are not in topological order"; return GB_ERR_ARG; }
        }
        // node lengths live on the device; bound the traceback by the host copy of the index sizes instead:
        // every node is at most 1023 bases (position encoding), the exact sum is taken below from d->h_node_len
        for (uint64_t u = 0; u < N; u++) bases += d->h_node_len[node[node_off[p] + u]];
        col_off[p + 1] = col_off[p] + N * (uint64_t)(m + 1);
        tb_off[p + 1] = tb_off[p] + bases * (uint64_t)(m + 1);
        step_off[p + 1] = step_off[p] + bases + m + 2;
    }
    const uint32_t Lc = std::max<uint32_t>(32u, (max_q + 15u) & ~15u);
    if (Lc > 1024) { g_last_error = "query too long"; return GB_ERR_ARG; }
    if (tb_off[n] > (uint64_t)8 << 30) { g_last_error = "gb_sw_batch: batch needs more than 8 GiB of traceback"; return GB_ERR_CAPACITY; }
    DevBuf<uint32_t> d_node, d_pred, d_edits, d_nm, d_ne, d_counter, d_colstart, d_seqoff, d_seqlen, d_steps;
    DevBuf<uint64_t> d_noff, d_poff, d_qoff, d_coloff, d_tboff, d_stepoff;
    DevBuf<int32_t> d_score, d_lastH, d_lastE; DevBuf<uint8_t> d_q, d_status, d_argH, d_argE, d_tb; DevBuf<gb_mapping> d_maps;
    int rc;
    if ((rc = d_node.upload(node, total_nodes ? total_nodes : 1, d->stream, total_nodes))) return rc;
    if ((rc = d_pred.upload(pred, total_pred ? total_pred : 1, d->stream, total_pred))) return rc;
    if ((rc = d_noff.upload(node_off, n + 1, d->stream))) return rc;
    if ((rc = d_poff.upload(pred_off, total_nodes + 1, d->stream))) return rc;
    if ((rc = d_q.upload(query, total_q ? total_q : 1, d->stream, total_q))) return rc;
    if ((rc = d_qoff.upload(query_off, n + 1, d->stream))) return rc;
    if ((rc = d_coloff.upload(col_off.data(), n + 1, d->stream))) return rc;
    if ((rc = d_tboff.upload(tb_off.data(), n + 1, d->stream))) return rc;
    if ((rc = d_stepoff.upload(step_off.data(), n + 1, d->stream))) return rc;
    if ((rc = d_score.reserve(n)) || (rc = d_nm.reserve(n)) || (rc = d_ne.reserve(n)) || (rc = d_status.reserve(n))) return rc;
    if ((rc = d_maps.reserve((size_t)n * map_cap)) || (rc = d_edits.reserve((size_t)n * edit_cap))) return rc;
    if ((rc = d_lastH.reserve(col_off[n] + 1)) || (rc = d_lastE.reserve(col_off[n] + 1))) return rc;
    if ((rc = d_argH.reserve(col_off[n] + 1)) || (rc = d_argE.reserve(col_off[n] + 1))) return rc;
    if ((rc = d_tb.reserve(tb_off[n] + 1)) || (rc = d_steps.reserve(step_off[n] + 1))) return rc;
    if ((rc = d_colstart.reserve(total_nodes + 1)) || (rc = d_seqoff.reserve(total_nodes + 1)) || (rc = d_seqlen.reserve(total_nodes + 1))) return rc;
    if ((rc = d_counter.reserve(4))) return rc;
    GB_CUDA(cudaMemsetAsync(d_counter.ptr, 0, 16, d->stream));
    const uint32_t W = Lc + 1;
    const size_t per_warp = (((size_t)Lc + 15) & ~(size_t)15) + (size_t)W * 4 * 4 + 64;
    const size_t smem = per_warp * SW_WARPS;
    if (smem > 48 * 1024) GB_CUDA(cudaFuncSetAttribute(sw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int bps = 0;
    GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, sw_kernel, SW_WARPS * 32, smem));
    if (bps < 1) bps = 1;
    const uint32_t grid = std::max<uint32_t>(1u, std::min<uint32_t>((uint32_t)d->n_sms * bps, (n + SW_WARPS - 1) / SW_WARPS));
    SwBatch b;
    b.node = d_node.ptr; b.node_off = d_noff.ptr; b.pred = d_pred.ptr; b.pred_off = d_poff.ptr; b.query = d_q.ptr; b.query_off = d_qoff.ptr;
    b.n = n; b.Lc = Lc; b.score = d_score.ptr; b.maps = d_maps.ptr; b.edits = d_edits.ptr; b.n_maps = d_nm.ptr; b.n_edits = d_ne.ptr;
    b.status = d_status.ptr; b.map_cap = map_cap; b.edit_cap = edit_cap;
    b.lastH = d_lastH.ptr; b.lastE = d_lastE.ptr; b.argH = d_argH.ptr; b.argE = d_argE.ptr; b.col_off = d_coloff.ptr;
    b.tb = d_tb.ptr; b.tb_off = d_tboff.ptr; b.col_start = d_colstart.ptr; b.seq_off = d_seqoff.ptr; b.seq_len = d_seqlen.ptr;
    b.steps = d_steps.ptr; b.step_off = d_stepoff.ptr; b.work_counter = d_counter.ptr;
    GB_CUDA(cudaEventRecord(d->ev0, d->stream));
    sw_kernel<<<grid, SW_WARPS * 32, smem, d->stream>>>(d->ix, d->sc, b);
    d->launches++;
    GB_CUDA(cudaGetLastError());
    GB_CUDA(cudaEventRecord(d->ev1, d->stream));
    GB_CUDA(cudaMemcpyAsync(score, d_score.ptr, 4 * (size_t)n, cudaMemcpyDeviceToHost, d->stream));
    GB_CUDA(cudaMemcpyAsync(n_maps, d_nm.ptr, 4 * (size_t)n, cudaMemcpyDeviceToHost, d->stream));
    GB_CUDA(cudaMemcpyAsync(n_edits, d_ne.ptr, 4 * (size_t)n, cudaMemcpyDeviceToHost, d->stream));
    GB_CUDA(cudaMemcpyAsync(status, d_status.ptr, n, cudaMemcpyDeviceToHost, d->stream));
    GB_CUDA(cudaMemcpyAsync(maps, d_maps.ptr, sizeof(gb_mapping) * (size_t)n * map_cap, cudaMemcpyDeviceToHost, d->stream));
    GB_CUDA(cudaMemcpyAsync(edits, d_edits.ptr, 4 * (size_t)n * edit_cap, cudaMemcpyDeviceToHost, d->stream));
    GB_CUDA(cudaStreamSynchronize(d->stream));
    GB_CUDA(cudaEventElapsedTime(&d->last_kernel_ms, d->ev0, d->ev1));
    return GB_OK;
}
