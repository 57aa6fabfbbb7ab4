// xdrop_dag.cu — B3 stage seam: Aligner::align_xdrop(alignment, graph, order, mems, rc = false, max_gap)
// (aligner.cpp:833-855 -> DozeuInterface::align, dozeu_interface.cpp:608-685), the seeded two-pass
// X-drop alignment vg giraffe uses for mate rescue (minimizer_mapper.cpp:3385):
//   pass 1  from the seed, extend the query suffix to the right and take the best end as the "head"
//           (no seed: scan the last 15 query bases over the subgraph for their best local match);
//   pass 2  from the head, extend the query prefix to the LEFT with traceback; the query right of the
//           head is a soft clip.
// Cell rules: those of the pinned aligner (tail.cuh) extended to DAGs as oracle/xdrop_dag.cpp states.
//
// One warp per problem.  Pass 2 runs on the mirrored problem without materialising it: pass order t
// maps to node N - 1 - t, pass column c to base len - 1 - c, predecessors are the successors.
#include <cub/cub.cuh>
#include "giraffe_b200.h"
#include "device_state.cuh"
#include "align.cuh"

#include <algorithm>
#include <vector>

namespace gb {

constexpr int XD_WARPS = 4;

struct XdNode { int32_t lineage_max; uint32_t first_col; uint8_t computed, live, band_lo, band_hi; };

struct XdBatch {
    const uint32_t* node; const uint64_t* node_off;
    const uint32_t* pred; const uint64_t* pred_off;         // CSR over global node index (forward predecessors, ascending)
    const uint32_t* succ; const uint64_t* succ_off;         // forward successors, ascending
    const uint8_t* query; const uint64_t* query_off;
    const uint32_t* seed;                                   // 3 per problem: node index (0xffffffff: none), node offset, query offset
    const uint32_t* max_gap;
    uint32_t n; uint32_t Lc;
    int32_t* score; gb_mapping* maps; uint32_t* edits; uint32_t* n_maps; uint32_t* n_edits; uint8_t* status;
    uint32_t map_cap, edit_cap;
    int32_t* lastH; int32_t* lastE; uint8_t* argH; uint8_t* argE; const uint64_t* col_off;
    uint8_t* tb; const uint64_t* tb_off;
    uint32_t* col_start; uint32_t* seq_off; uint32_t* seq_len; XdNode* nstate;
    uint32_t* steps; const uint64_t* step_off;
    uint32_t* work_counter;
};

struct XdBest { int32_t best; uint32_t t, c, j; };

// One pinned X-drop pass over the (possibly mirrored) problem, starting before pass column `o` of pass
// node `s`, query q[0..m) in shared memory.  TB: write traceback bytes.
template <bool MIRROR, bool TB>
__device__ inline XdBest xd_pass(const DevIndex& ix, const DevScores& sc, const XdBatch& b, uint32_t p, uint32_t N, uint64_t n0,
                                 const uint8_t* q, uint32_t m, uint32_t s, uint32_t o, uint32_t max_gap, DpSmem dps) {
    const int lane = lane_id();
    const uint32_t W = m + 1, n_chunks = (W + 31) >> 5;
    const int32_t go = sc.gap_open, ge = sc.gap_extend;
    const int32_t xt = go + ge * ((int32_t)max_gap - 1);
    int32_t* lastH = b.lastH + b.col_off[p]; int32_t* lastE = b.lastE + b.col_off[p];
    uint8_t* argH = b.argH + b.col_off[p]; uint8_t* argE = b.argE + b.col_off[p];
    uint8_t* tb = b.tb + b.tb_off[p];
    XdNode* ns = b.nstate + n0;
    // the workspace rows are addressed by pass node index t with this pass's W
    for (uint32_t t = lane; t < N; t += 32) { XdNode z; z.lineage_max = 0; z.first_col = 0; z.computed = 0; z.live = 0; z.band_lo = 0; z.band_hi = 0; ns[t] = z; }
    __syncwarp();
    int32_t lane_best = DP_NEG; uint32_t lane_t = 0, lane_c = 0, lane_j = 0;
    for (uint32_t t = s; t < N; t++) {
        const uint32_t u = MIRROR ? N - 1 - t : t;
        const uint32_t len = b.seq_len[n0 + u], soff = b.seq_off[n0 + u];
        uint32_t plo, phi, first_col = 0; int32_t run_max;
        if (t == s) {
            for (uint32_t j = lane; j < W; j += 32) {
                int32_t h = DP_NEG;
                if (j == 0) h = 0; else if (j <= max_gap) h = -(go + (int32_t)(j - 1) * ge);
                dps.Hp[j] = h; dps.Ep[j] = DP_NEG;
            }
            plo = 0; phi = min(n_chunks, (min(max_gap, m) >> 5) + 1);
            run_max = 0; first_col = o;
        } else {
            // merge the last columns of the computed, live predecessors (pass order)
            const uint64_t e0 = MIRROR ? b.succ_off[n0 + u] : b.pred_off[n0 + u];
            const uint32_t ne = (uint32_t)((MIRROR ? b.succ_off[n0 + u + 1] : b.pred_off[n0 + u + 1]) - e0);
            plo = n_chunks; phi = 0; run_max = INT_MIN;
            bool any = false;
            for (uint32_t pi = 0; pi < ne; pi++) {
                const uint32_t pu = MIRROR ? b.succ[e0 + pi] : b.pred[e0 + pi];
                const uint32_t pt = MIRROR ? N - 1 - pu : pu;
                if (pt < s) continue;
                const XdNode pn = ns[pt];
                if (!pn.computed || !pn.live) continue;
                any = true;
                plo = min(plo, (uint32_t)pn.band_lo); phi = max(phi, (uint32_t)pn.band_hi);
                run_max = max(run_max, pn.lineage_max);
            }
            if (!any) continue;
            for (uint32_t j = plo * 32 + lane; j < min(W, phi * 32); j += 32) {
                int32_t h = DP_NEG, e = DP_NEG; uint32_t ah = 0xff, ae = 0xff;
                for (uint32_t pi = 0; pi < ne; pi++) {
                    const uint32_t pu = MIRROR ? b.succ[e0 + pi] : b.pred[e0 + pi];
                    const uint32_t pt = MIRROR ? N - 1 - pu : pu;
                    if (pt < s) continue;
                    const XdNode pn = ns[pt];
                    if (!pn.computed || !pn.live) continue;
                    if ((j >> 5) < pn.band_lo || (j >> 5) >= pn.band_hi) continue;
                    const int32_t ph = lastH[(size_t)pt * W + j], pe = lastE[(size_t)pt * W + j];
                    if (ph > h) { h = ph; ah = pi; }
                    if (pe > e) { e = pe; ae = pi; }
                }
                dps.Hp[j] = h; dps.Ep[j] = e;
                if (TB) { argH[(size_t)t * W + j] = (uint8_t)ah; argE[(size_t)t * W + j] = (uint8_t)ae; }
            }
        }
        __syncwarp();
        for (uint32_t c = first_col; c < len && plo < phi; c++) {
            const uint8_t r = __ldg(ix.seq + soff + (MIRROR ? len - 1 - c : c));
            uint8_t* tbcol = tb + (size_t)(b.col_start[n0 + u] + c) * W;
            int32_t carry = INT_MIN, prevH_last = DP_NEG, prev_ph_last = DP_NEG, col_max = DP_NEG;
            uint32_t clo = n_chunks, chi = 0;
            for (uint32_t ch = plo; ch < n_chunks; ch++) {
                const uint32_t j = ch * 32 + lane;
                const bool in = j < W;
                const bool pin = in && ch < phi;
                int32_t ph = DP_NEG, pe = DP_NEG;
                if (pin) { ph = dps.Hp[j]; pe = dps.Ep[j]; }
                int32_t phm1 = __shfl_up_sync(FULL, ph, 1);
                if (lane == 0) phm1 = prev_ph_last;
                prev_ph_last = __shfl_sync(FULL, ph, 31);
                int32_t e = DP_NEG;
                if (ph > DP_NEG) e = ph - go;
                if (pe > DP_NEG) e = max(e, pe - ge);
                int32_t d = DP_NEG;
                if (in && j > 0 && phm1 > DP_NEG) {
                    const uint8_t qc = q[j - 1];
                    int32_t sub = (qc == r) ? sc.match : -sc.mismatch;
                    if (j == m) sub += sc.full_length_bonus;
                    d = phm1 + sub;
                }
                const int32_t hprime = max(d, e);
                int32_t g = (in && hprime > DP_NEG) ? hprime + (int32_t)j * ge : INT_MIN;
                int32_t incl = g;
#pragma unroll
                for (int ofs = 1; ofs < 32; ofs <<= 1) { const int32_t tt = __shfl_up_sync(FULL, incl, ofs); if (lane >= ofs) incl = max(incl, tt); }
                int32_t excl = __shfl_up_sync(FULL, incl, 1);
                if (lane == 0) excl = INT_MIN;
                excl = max(excl, carry);
                carry = max(carry, __shfl_sync(FULL, incl, 31));
                int32_t f = DP_NEG;
                if (in && j > 0 && excl > INT_MIN) f = excl - go - (int32_t)(j - 1) * ge;
                int32_t h = max(hprime, f);
                int32_t hm1 = __shfl_up_sync(FULL, h, 1);
                if (lane == 0) hm1 = prevH_last;
                prevH_last = __shfl_sync(FULL, h, 31);
                uint8_t tbv = 0;
                if (TB && in) {
                    if (d > DP_NEG && d == h) tbv = 0; else if (e > DP_NEG && e == h) tbv = 1; else tbv = 2;
                    if (ph > DP_NEG && e == ph - go) tbv |= 4;
                    if (j > 0 && hm1 > DP_NEG && f == hm1 - go) tbv |= 8;
                }
                if (in && h > DP_NEG && h < run_max - xt) { h = DP_NEG; e = DP_NEG; }
                if (in) { dps.Hc[j] = h; dps.Ec[j] = e; if (TB) tbcol[j] = tbv; }
                const bool alive = in && h > DP_NEG;
                if (alive) {
                    col_max = max(col_max, h);
                    if (h > lane_best) { lane_best = h; lane_t = t; lane_c = c; lane_j = j; }
                }
                if (__any_sync(FULL, alive)) { clo = min(clo, ch); chi = ch + 1; }
                else if (ch >= phi) break;
            }
            __syncwarp();
            col_max = __reduce_max_sync(FULL, col_max);
            if (col_max > run_max) run_max = col_max;
            int32_t* t1p = dps.Hp; dps.Hp = dps.Hc; dps.Hc = t1p;
            int32_t* t2p = dps.Ep; dps.Ep = dps.Ec; dps.Ec = t2p;
            plo = clo; phi = chi;
        }
        if (plo < phi) for (uint32_t j = plo * 32 + lane; j < min(W, phi * 32); j += 32) { lastH[(size_t)t * W + j] = dps.Hp[j]; lastE[(size_t)t * W + j] = dps.Ep[j]; }
        if (lane == 0) {
            XdNode z; z.lineage_max = run_max; z.first_col = first_col; z.computed = 1; z.live = plo < phi ? 1 : 0;
            z.band_lo = (uint8_t)min(plo, 255u); z.band_hi = (uint8_t)(plo < phi ? phi : min(plo, 255u));
            ns[t] = z;
        }
        __syncwarp();
    }
    XdBest r; r.best = lane_best; r.t = lane_t; r.c = lane_c; r.j = lane_j;
#pragma unroll
    for (int ofs = 16; ofs > 0; ofs >>= 1) {
        const int32_t ob = __shfl_xor_sync(FULL, r.best, ofs);
        const uint32_t ot = __shfl_xor_sync(FULL, r.t, ofs), oc = __shfl_xor_sync(FULL, r.c, ofs), oj = __shfl_xor_sync(FULL, r.j, ofs);
        const bool take = ob > r.best || (ob == r.best && (ot < r.t || (ot == r.t && (oc < r.c || (oc == r.c && oj < r.j)))));
        if (take) { r.best = ob; r.t = ot; r.c = oc; r.j = oj; }
    }
    return r;
}

// Local scan of a short query (bonus at its right end only): best end cell; columns in shared memory,
// last columns per node in the workspace (oracle/xdrop_dag.cpp scan_local).
__device__ inline XdBest xd_scan(const DevIndex& ix, const DevScores& sc, const XdBatch& b, uint32_t p, uint32_t N, uint64_t n0,
                                 const uint8_t* q, uint32_t m, DpSmem dps) {
    const int lane = lane_id();
    const uint32_t W = m + 1;         // m <= 15: a single chunk
    const int32_t go = sc.gap_open, ge = sc.gap_extend, bonus = sc.full_length_bonus;
    int32_t* lastH = b.lastH + b.col_off[p]; int32_t* lastE = b.lastE + b.col_off[p];
    int32_t lane_best = 0; uint32_t lane_t = 0, lane_c = 0, lane_j = 0;
    const uint32_t j = lane; const bool in = j < W;
    for (uint32_t u = 0; u < N; u++) {
        const uint32_t len = b.seq_len[n0 + u], soff = b.seq_off[n0 + u];
        int32_t ph = DP_NEG, pe = DP_NEG;
        const uint64_t e0 = b.pred_off[n0 + u]; const uint32_t ne = (uint32_t)(b.pred_off[n0 + u + 1] - e0);
        if (in) for (uint32_t pi = 0; pi < ne; pi++) { const uint32_t pu = b.pred[e0 + pi]; ph = max(ph, lastH[(size_t)pu * W + j]); pe = max(pe, lastE[(size_t)pu * W + j]); }
        for (uint32_t c = 0; c < len; c++) {
            const uint8_t r = __ldg(ix.seq + soff + c);
            const int32_t phm1 = __shfl_up_sync(FULL, ph, 1);
            int32_t d = DP_NEG, e = DP_NEG;
            if (in && j > 0) {
                const uint8_t qc = q[j - 1];
                const int32_t sub = (!is_acgt(r) || !is_acgt(qc)) ? 0 : (qc == r ? sc.match : -sc.mismatch);
                d = max(phm1, 0) + sub;
                if (ph > 0) e = ph - go;
                if (pe > DP_NEG) e = max(e, pe - ge);
                if (e <= 0) e = DP_NEG;
            }
            const int32_t hprime = max(d, e);
            int32_t incl = (in && j > 0 && hprime > 0) ? hprime + (int32_t)j * ge : INT_MIN;
#pragma unroll
            for (int ofs = 1; ofs < 32; ofs <<= 1) { const int32_t tt = __shfl_up_sync(FULL, incl, ofs); if (lane >= ofs) incl = max(incl, tt); }
            int32_t excl = __shfl_up_sync(FULL, incl, 1);
            if (lane == 0) excl = INT_MIN;
            int32_t f = DP_NEG;
            if (in && j > 0 && excl > INT_MIN) f = excl - go - (int32_t)(j - 1) * ge;
            if (f <= 0) f = DP_NEG;
            int32_t h = 0;
            if (in && j > 0) h = max(max(d, 0), max(e, f));
            if (in && j > 0) {
                int32_t cand = h;
                if (j == m && d + bonus >= h) cand = d + bonus;
                if (cand > lane_best) { lane_best = cand; lane_t = u; lane_c = c; lane_j = j; }
            }
            ph = in ? h : DP_NEG; pe = e;
        }
        if (in) { lastH[(size_t)u * W + j] = ph; lastE[(size_t)u * W + j] = pe; }
        __syncwarp();
    }
    XdBest r; r.best = lane_best; r.t = lane_t; r.c = lane_c; r.j = lane_j;
#pragma unroll
    for (int ofs = 16; ofs > 0; ofs >>= 1) {
        const int32_t ob = __shfl_xor_sync(FULL, r.best, ofs);
        const uint32_t ot = __shfl_xor_sync(FULL, r.t, ofs), oc = __shfl_xor_sync(FULL, r.c, ofs), oj = __shfl_xor_sync(FULL, r.j, ofs);
        const bool take = ob > r.best || (ob == r.best && (ot < r.t || (ot == r.t && (oc < r.c || (oc == r.c && oj < r.j)))));
        if (take) { r.best = ob; r.t = ot; r.c = oc; r.j = oj; }
    }
    (void)dps;
    return r;
}

__device__ inline void xd_problem(const DevIndex& ix, const DevScores& sc, const XdBatch& b, uint32_t p,
                                  uint8_t* q, DpSmem dps, int32_t& score_out, PathBuf& out, uint32_t& status) {
    const int lane = lane_id();
    const uint64_t n0 = b.node_off[p]; const uint32_t N = (uint32_t)(b.node_off[p + 1] - n0);
    const uint64_t q0 = b.query_off[p]; const uint32_t m = (uint32_t)(b.query_off[p + 1] - q0);
    score_out = 0; pb_reset(out);
    if (N == 0 || m == 0) return;
    const uint32_t max_gap = max(b.max_gap[p], 1u);
    if (lane == 0) {
        uint32_t cols = 0;
        for (uint32_t u = 0; u < N; u++) {
            const gb_node_rec nr = load_node(ix, b.node[n0 + u]);
            b.seq_off[n0 + u] = nr.seq_off; b.seq_len[n0 + u] = nr.len; b.col_start[n0 + u] = cols;
            cols += nr.len;
        }
    }
    __syncwarp();
    const uint32_t seed_u = b.seed[3 * p], seed_o = b.seed[3 * p + 1], seed_q = b.seed[3 * p + 2];
    // ---- pass 1: the head ---------------------------------------------------------------------------
    uint32_t head_u, head_off, head_q;
    if (seed_u != 0xffffffffu) {
        if (seed_u >= N || seed_q > m || seed_o > b.seq_len[n0 + seed_u]) { status = GB_ITEM_OUT_FULL; return; }
        const uint32_t m1 = m - seed_q;
        for (uint32_t i = lane; i < m1; i += 32) q[i] = dp_query_base(b.query[q0 + seed_q + i]);
        __syncwarp();
        const XdBest r = xd_pass<false, false>(ix, sc, b, p, N, n0, q, m1, seed_u, seed_o, max_gap, dps);
        if (r.best > 0) { head_u = r.t; head_off = r.c + 1; head_q = seed_q + r.j; }
        else { head_u = seed_u; head_off = seed_o; head_q = seed_q; }
    } else {
        const uint32_t scan_len = min(m, 15u);
        for (uint32_t i = lane; i < scan_len; i += 32) q[i] = b.query[q0 + m - scan_len + i];
        __syncwarp();
        const XdBest r = xd_scan(ix, sc, b, p, N, n0, q, scan_len, dps);
        if (r.best <= 0) return;
        head_u = r.t; head_off = r.c + 1; head_q = (m - scan_len) + r.j;
    }
    __syncwarp();
    if (head_q == 0) return;
    // ---- pass 2: leftwards from the head (mirrored problem) -------------------------------------------
    const uint32_t m2 = head_q;
    for (uint32_t i = lane; i < m2; i += 32) q[i] = dp_query_base(b.query[q0 + head_q - 1 - i]);
    __syncwarp();
    const uint32_t rs = N - 1 - head_u, ro = b.seq_len[n0 + head_u] - head_off;
    const XdBest r2 = xd_pass<true, true>(ix, sc, b, p, N, n0, q, m2, rs, ro, max_gap, dps);
    if (r2.best <= 0) return;
    score_out = r2.best;

    // ---- traceback in mirrored space: recorded end -> start = forward alignment start -> head ------------
    const uint32_t W = m2 + 1;
    const uint8_t* argH = b.argH + b.col_off[p]; const uint8_t* argE = b.argE + b.col_off[p];
    const uint8_t* tb = b.tb + b.tb_off[p];
    const XdNode* ns = b.nstate + n0;
    uint32_t* steps = b.steps + b.step_off[p];
    uint32_t n_steps = 0;
    {
        uint32_t t = r2.t, c = r2.c, j = r2.j;
        int state = 0; bool at_virtual = false;
        while (true) {
            if (at_virtual) { for (; j > 0; j--) { if (lane == 0) steps[n_steps] = (rs << 8) | 2u; n_steps++; } break; }
            const uint32_t u = N - 1 - t;
            const uint8_t tbv = tb[(size_t)(b.col_start[n0 + u] + c) * W + j];
            const bool first = c == ns[t].first_col;
            // previous column for query offset jj, through E or not
            auto go_prev = [&](uint32_t jj, bool via_E) {
                if (!first) { c--; return; }
                if (t == rs) { at_virtual = true; return; }
                uint32_t tt = t;
                while (true) {
                    const uint32_t uu = N - 1 - tt;
                    const uint32_t pi = via_E ? argE[(size_t)tt * W + jj] : argH[(size_t)tt * W + jj];
                    const uint32_t pu = b.succ[b.succ_off[n0 + uu] + pi];
                    const uint32_t pt = N - 1 - pu;
                    if (b.seq_len[n0 + pu] > ns[pt].first_col) { t = pt; c = b.seq_len[n0 + pu] - 1; return; }
                    if (pt == rs) { t = rs; at_virtual = true; return; }
                    tt = pt;
                }
            };
            if (state == 0) {
                const uint32_t src = tbv & 3u;
                if (src == 0) {
                    const uint8_t qc = q[j - 1], r = __ldg(ix.seq + b.seq_off[n0 + u] + (b.seq_len[n0 + u] - 1 - c));
                    if (lane == 0) steps[n_steps] = (t << 8) | (qc == r ? 0u : 1u);
                    n_steps++;
                    j--; go_prev(j, false);
                    if (at_virtual && j == 0) break;
                    continue;
                }
                state = src == 1 ? 1 : 2;
                continue;
            }
            if (state == 1) {
                if (lane == 0) steps[n_steps] = (t << 8) | 3u;
                n_steps++;
                const bool open = (tbv & 4u) != 0;
                go_prev(j, !open);
                state = open ? 0 : 1;
                if (at_virtual && j == 0 && state == 0) break;
                continue;
            }
            if (lane == 0) steps[n_steps] = (t << 8) | 2u;
            n_steps++;
            const bool open = (tbv & 8u) != 0;
            j--;
            state = open ? 0 : 2;
        }
    }
    __syncwarp();

    // ---- steps -> mappings in forward order -------------------------------------------------------------
    if (lane == 0) {
        uint32_t aligned_q = 0;
        for (uint32_t i = 0; i < n_steps; i++) aligned_q += (steps[i] & 0xffu) != 3u ? 1u : 0u;
        uint32_t qpos = head_q - aligned_q;
        const uint8_t* qf = b.query + q0;           // forward query (raw bytes) for substitution bases
        uint32_t i = 0;
        while (i < n_steps) {
            const uint32_t t = steps[i] >> 8, u = N - 1 - t;
            uint32_t k = i, cols = 0;
            while (k < n_steps && (steps[k] >> 8) == t) { if ((steps[k] & 0xffu) != 2u) cols++; k++; }
            const uint32_t end = (u == head_u) ? head_off : b.seq_len[n0 + u];
            pb_add_mapping(out, u, end - cols);
            if (i == 0 && qpos > 0) pb_add_edit(out, edit_word(GB_EDIT_INS, qpos, 0));
            uint32_t cur = 0xff, run = 0;
            auto flush = [&]() {
                if (cur == 1) { for (uint32_t x = 0; x < run; x++) { pb_add_edit(out, edit_word(GB_EDIT_SUB, 1, base2(qf[qpos]))); qpos++; } }
                else if (run > 0) {
                    if (cur == 0) { pb_add_edit(out, edit_word(GB_EDIT_MATCH, run, 0)); qpos += run; }
                    else if (cur == 2) {
                        // an insertion next to an insertion edit of this mapping (the leading soft clip) extends it
                        if (!out.overflow && out.maps[out.n_maps - 1].n_edits > 0 && (out.edits[out.n_edits - 1] & 3u) == GB_EDIT_INS) out.edits[out.n_edits - 1] += run << 4;
                        else pb_add_edit(out, edit_word(GB_EDIT_INS, run, 0));
                        qpos += run;
                    }
                    else pb_add_edit(out, edit_word(GB_EDIT_DEL, run, 0));
                }
            };
            for (uint32_t x = i; x < k; x++) {
                const uint32_t op = steps[x] & 0xffu;
                if (op == cur) run++; else { if (cur != 0xff) flush(); cur = op; run = 1; }
            }
            if (cur != 0xff) flush();
            i = k;
        }
        if (out.n_maps > 0 && head_q < m && !out.overflow) {
            if (out.maps[out.n_maps - 1].n_edits > 0 && (out.edits[out.n_edits - 1] & 3u) == GB_EDIT_INS) out.edits[out.n_edits - 1] += (m - head_q) << 4;
            else pb_add_edit(out, edit_word(GB_EDIT_INS, m - head_q, 0));
        }
    }
    __syncwarp();
    out.n_maps = __shfl_sync(FULL, out.n_maps, 0); out.n_edits = __shfl_sync(FULL, out.n_edits, 0);
    out.overflow = __shfl_sync(FULL, (int)out.overflow, 0) != 0;
    if (out.overflow) status = GB_ITEM_OUT_FULL;
}

__global__ void __launch_bounds__(XD_WARPS * 32)
xdrop_dag_kernel(DevIndex ix, DevScores sc, XdBatch b) {
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
        xd_problem(ix, sc, b, p, q, dps, score, out, status);
        if (status != GB_ITEM_OK || score <= 0) { score = 0; out.n_maps = 0; out.n_edits = 0; }
        if (lane == 0) { b.score[p] = score; b.n_maps[p] = out.n_maps; b.n_edits[p] = out.n_edits; b.status[p] = (uint8_t)status; }
        __syncwarp();
    }
}

} // namespace gb

using namespace gb;

extern "C" int gb_xdrop_dag_batch(gb_device* d, uint32_t n,
                                  const uint32_t* node, const uint64_t* node_off, const uint32_t* pred, const uint64_t* pred_off,
                                  const uint8_t* query, const uint64_t* query_off, const uint32_t* seed, const uint32_t* max_gap,
                                  uint32_t map_cap, uint32_t edit_cap,
                                  int32_t* score, gb_mapping* maps, uint32_t* edits, uint32_t* n_maps, uint32_t* n_edits, uint8_t* status) {
    if (!d || !node || !node_off || !pred || !pred_off || !query || !query_off || !seed || !max_gap || !score || !maps || !edits ||
        !n_maps || !n_edits || !status || map_cap == 0 || edit_cap == 0) return GB_ERR_ARG;
    if (n == 0) return GB_OK;
    GB_CUDA(cudaSetDevice(d->device));
    const uint64_t total_nodes = node_off[n], total_pred = pred_off[total_nodes], total_q = query_off[n];
    uint32_t max_q = 0;
    std::vector<uint64_t> col_off(n + 1, 0), tb_off(n + 1, 0), step_off(n + 1, 0), succ_off(total_nodes + 1, 0);
    std::vector<uint32_t> succ(total_pred ? total_pred : 1);
    for (uint32_t p = 0; p < n; p++) {
        const uint32_t m = (uint32_t)(query_off[p + 1] - query_off[p]);
        max_q = std::max(max_q, m);
        uint64_t bases = 0;
        const uint64_t n0 = node_off[p], N = node_off[p + 1] - n0;
        std::vector<uint32_t> deg(N, 0);
        for (uint64_t u = 0; u < N; u++) {
            const uint32_t v = node[n0 + u];
            if (v < 2 || v >= d->ix.n_nodes) { g_last_error = "gb_xdrop_dag_batch: unknown node"; return GB_ERR_ARG; }
            if (pred_off[n0 + u + 1] - pred_off[n0 + u] > 255) { g_last_error = "gb_xdrop_dag_batch: more than 255 predecessors"; return GB_ERR_ARG; }
            for (uint64_t e = pred_off[n0 + u]; e < pred_off[n0 + u + 1]; e++) {
                if (pred[e] >= u) { g_last_error = "gb_xdrop_dag_batch: nodes are not in topological order"; return GB_ERR_ARG; }
                deg[pred[e]]++;
            }
            bases += d->h_node_len[v];
        }
        // successor CSR, ascending (nodes are visited in ascending order)
        for (uint64_t u = 0; u < N; u++) { if (deg[u] > 255) { g_last_error = "gb_xdrop_dag_batch: more than 255 successors"; return GB_ERR_ARG; } succ_off[n0 + u + 1] = succ_off[n0 + u] + deg[u]; }
        std::vector<uint32_t> fill(N, 0);
        for (uint64_t u = 0; u < N; u++)
            for (uint64_t e = pred_off[n0 + u]; e < pred_off[n0 + u + 1]; e++) { const uint32_t pu = pred[e]; succ[succ_off[n0 + pu] + fill[pu]++] = (uint32_t)u; }
        col_off[p + 1] = col_off[p] + N * (uint64_t)(m + 1);
        tb_off[p + 1] = tb_off[p] + bases * (uint64_t)(m + 1);
        step_off[p + 1] = step_off[p] + bases + m + 2;
    }
    const uint32_t Lc = std::max<uint32_t>(32u, (max_q + 15u) & ~15u);
    if (Lc > 1024) { g_last_error = "query too long"; return GB_ERR_ARG; }
    if (tb_off[n] > (uint64_t)8 << 30) { g_last_error = "gb_xdrop_dag_batch: batch needs more than 8 GiB of traceback"; return GB_ERR_CAPACITY; }
    DevBuf<uint32_t> d_node, d_pred, d_succ, d_seed, d_gap, d_edits, d_nm, d_ne, d_counter, d_colstart, d_seqoff, d_seqlen, d_steps;
    DevBuf<uint64_t> d_noff, d_poff, d_soff, d_qoff, d_coloff, d_tboff, d_stepoff;
    DevBuf<int32_t> d_score, d_lastH, d_lastE; DevBuf<uint8_t> d_q, d_status, d_argH, d_argE, d_tb; DevBuf<gb_mapping> d_maps; DevBuf<XdNode> d_nstate;
    int rc;
    if ((rc = d_node.upload(node, total_nodes ? total_nodes : 1, d->stream, total_nodes))) return rc;
    if ((rc = d_pred.upload(pred, total_pred ? total_pred : 1, d->stream, total_pred))) return rc;
    if ((rc = d_succ.upload(succ.data(), total_pred ? total_pred : 1, d->stream, total_pred))) return rc;
    if ((rc = d_noff.upload(node_off, n + 1, d->stream))) return rc;
    if ((rc = d_poff.upload(pred_off, total_nodes + 1, d->stream))) return rc;
    if ((rc = d_soff.upload(succ_off.data(), total_nodes + 1, d->stream))) return rc;
    if ((rc = d_q.upload(query, total_q ? total_q : 1, d->stream, total_q))) return rc;
    if ((rc = d_qoff.upload(query_off, n + 1, d->stream))) return rc;
    if ((rc = d_seed.upload(seed, 3 * (size_t)n, d->stream))) return rc;
    if ((rc = d_gap.upload(max_gap, n, d->stream))) return rc;
    if ((rc = d_coloff.upload(col_off.data(), n + 1, d->stream))) return rc;
    if ((rc = d_tboff.upload(tb_off.data(), n + 1, d->stream))) return rc;
    if ((rc = d_stepoff.upload(step_off.data(), n + 1, d->stream))) return rc;
    if ((rc = d_score.reserve(n)) || (rc = d_nm.reserve(n)) || (rc = d_ne.reserve(n)) || (rc = d_status.reserve(n))) return rc;
    if ((rc = d_maps.reserve((size_t)n * map_cap)) || (rc = d_edits.reserve((size_t)n * edit_cap))) return rc;
    if ((rc = d_lastH.reserve(col_off[n] + 1)) || (rc = d_lastE.reserve(col_off[n] + 1))) return rc;
    if ((rc = d_argH.reserve(col_off[n] + 1)) || (rc = d_argE.reserve(col_off[n] + 1))) return rc;
    if ((rc = d_tb.reserve(tb_off[n] + 1)) || (rc = d_steps.reserve(step_off[n] + 1))) return rc;
    if ((rc = d_colstart.reserve(total_nodes + 1)) || (rc = d_seqoff.reserve(total_nodes + 1)) || (rc = d_seqlen.reserve(total_nodes + 1))) return rc;
    if ((rc = d_nstate.reserve(total_nodes + 1))) return rc;
    if ((rc = d_counter.reserve(4))) return rc;
    GB_CUDA(cudaMemsetAsync(d_counter.ptr, 0, 16, d->stream));
    const uint32_t W = Lc + 1;
    const size_t per_warp = (((size_t)Lc + 15) & ~(size_t)15) + (size_t)W * 4 * 4 + 64;
    const size_t smem = per_warp * XD_WARPS;
    if (smem > 48 * 1024) GB_CUDA(cudaFuncSetAttribute(xdrop_dag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int bps = 0;
    GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, xdrop_dag_kernel, XD_WARPS * 32, smem));
    if (bps < 1) bps = 1;
    const uint32_t grid = std::max<uint32_t>(1u, std::min<uint32_t>((uint32_t)d->n_sms * bps, (n + XD_WARPS - 1) / XD_WARPS));
    XdBatch b;
    b.node = d_node.ptr; b.node_off = d_noff.ptr; b.pred = d_pred.ptr; b.pred_off = d_poff.ptr; b.succ = d_succ.ptr; b.succ_off = d_soff.ptr;
    b.query = d_q.ptr; b.query_off = d_qoff.ptr; b.seed = d_seed.ptr; b.max_gap = d_gap.ptr;
    b.n = n; b.Lc = Lc; b.score = d_score.ptr; b.maps = d_maps.ptr; b.edits = d_edits.ptr; b.n_maps = d_nm.ptr; b.n_edits = d_ne.ptr;
    b.status = d_status.ptr; b.map_cap = map_cap; b.edit_cap = edit_cap;
    b.lastH = d_lastH.ptr; b.lastE = d_lastE.ptr; b.argH = d_argH.ptr; b.argE = d_argE.ptr; b.col_off = d_coloff.ptr;
    b.tb = d_tb.ptr; b.tb_off = d_tboff.ptr; b.col_start = d_colstart.ptr; b.seq_off = d_seqoff.ptr; b.seq_len = d_seqlen.ptr; b.nstate = d_nstate.ptr;
    b.steps = d_steps.ptr; b.step_off = d_stepoff.ptr; b.work_counter = d_counter.ptr;
    GB_CUDA(cudaEventRecord(d->ev0, d->stream));
    xdrop_dag_kernel<<<grid, XD_WARPS * 32, smem, d->stream>>>(d->ix, d->sc, b);
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
