// ORACLE — TEST INFRASTRUCTURE ONLY.  Full (unbanded) local alignment of a read against a DAG:
//   Aligner::align(Alignment&, const HandleGraph&, const vector<handle_t>& topological_order)   aligner.cpp:571-626
//   Aligner::align_internal (pinned = false)                                                     aligner.cpp:344-564
//     -> gssw_graph_fill_pinned(..., gap_open, gap_extension, full_length_bonus, full_length_bonus, ...)   :399-402
//     -> gssw_graph_trace_back(...)                                                              :537-548
//     -> gssw_mapping_to_alignment                                                               :120-241
// used by vg giraffe as the rescue fallback (fix_dozeu_score, minimizer_mapper.cpp:3502-3517) and for
// --rescue-algorithm gssw (:3390).
//
// PARITY UNPINNED beyond the reference's unit vectors: the DP itself lives in vgteam/gssw @ 14b4d43
// (ABSENT from /root/reference).  Smith-Waterman scores are implementation independent, so the SCORE
// is pinned by the recurrence below plus the vectors of src/unittest/aligner.cpp:23-345 (bonus at
// both ends, single-base read, softclip vs attached ends, N handling); which of several co-optimal
// alignments gssw returns is not visible from vg, so this file DEFINES:
//   * scoring: match / mismatch from gb_scores; a pair with N (any non-ACGT byte) on either side scores 0
//     (vg's 5x5 score matrix, aligner.cpp:30-47 via GSSWAligner::score_matrix);
//     gap of length n costs gap_open + (n - 1) * gap_extend;
//     full_length_bonus is added when query base 1 is aligned by a match/mismatch, and again when
//     query base m is (left / right soft clips forfeit it), aligner.cpp:399-402;
//   * local alignment: any prefix / suffix of the query may be soft clipped at no cost; a score <= 0
//     means "unaligned" (empty path, score 0);
//   * the optimal end cell is the first maximum in (topological node order, node offset, query offset);
//   * traceback prefers diagonal > deletion > insertion, opening a gap over extending it, the first
//     predecessor (in the caller's list order) on node boundaries, and stops as soon as the running
//     prefix is not worth keeping (prefix score <= 0);
//   * soft clips are reported as insertion edits on the first / last mapping, like
//     gssw_mapping_to_alignment does (aligner.cpp:150-241).
#include "tail_align.hpp"
#include <climits>

namespace oracle {

static inline bool is_acgt(char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

// Recurrence (cell = graph base c of node u x query prefix j, 0 <= j <= m):
//   H(., 0) = 0; gap states are "dead" unless their value is > 0 (a non-positive prefix is never kept)
//   d(j) = (j == 1 ? bonus : max(H_prev(j - 1), 0)) + s(q_j, r)                 j >= 1
//   e(j) = max(H_prev(j) - go  [H_prev(j) > 0],  E_prev(j) - ge [E_prev(j) alive])   dead if <= 0
//   f(j) = max(H(j - 1) - go   [H(j - 1) > 0],   F(j - 1) - ge  [F(j - 1) alive])    dead if <= 0
//   H(j) = max(d, e, f, 0)
// "prev" is the previous base of the node, or for the first base the element-wise maximum over the
// predecessors' last columns (first predecessor on ties; H and E merged independently).
// End of the alignment: max over all cells of H, and of d(., m) + bonus at j = m (on a tie with H
// the attached end is taken); first maximum in (u, c, j) order.
LocalAlignmentResult sw_local_dag(const Graph& g, const gb_scores& sc, const DagProblem& P, const std::string& q, uint64_t* cells) {
    constexpr int32_t NEG = INT_MIN / 4;
    const size_t N = P.node.size(), m = q.size(), W = m + 1;
    const int32_t go = sc.gap_open, ge = sc.gap_extend, bonus = sc.full_length_bonus;
    LocalAlignmentResult out;
    if (N == 0 || m == 0) return out;
    auto subst = [&](char a, char b) -> int32_t {
        if (!is_acgt(a) || !is_acgt(b)) return 0;
        return a == b ? (int32_t)sc.match : -(int32_t)sc.mismatch;
    };
    struct NodeDP { std::string_view seq; std::vector<int32_t> H, E, F, D, inH, inE, argH, argE; };
    std::vector<NodeDP> dp(N);
    int32_t best = 0; size_t best_u = 0, best_c = 0, best_j = 0; bool best_end_diag = false;
    for (size_t u = 0; u < N; u++) {
        NodeDP& n = dp[u];
        n.seq = g.get_sequence_view(P.node[u]);
        const size_t len = n.seq.size();
        n.H.assign(len * W, 0); n.E.assign(len * W, NEG); n.F.assign(len * W, NEG); n.D.assign(len * W, NEG);
        n.inH.assign(W, NEG); n.inE.assign(W, NEG); n.argH.assign(W, -1); n.argE.assign(W, -1);
        for (size_t pi = 0; pi < P.pred[u].size(); pi++) {
            const NodeDP& p = dp[P.pred[u][pi]];
            const size_t pl = p.seq.size();
            for (size_t j = 0; j < W; j++) {
                const int32_t h = p.H[(pl - 1) * W + j], e = p.E[(pl - 1) * W + j];
                if (h > n.inH[j]) { n.inH[j] = h; n.argH[j] = (int32_t)pi; }
                if (e > n.inE[j]) { n.inE[j] = e; n.argE[j] = (int32_t)pi; }
            }
        }
        for (size_t c = 0; c < len; c++) {
            const int32_t* pH = c ? &n.H[(c - 1) * W] : n.inH.data();
            const int32_t* pE = c ? &n.E[(c - 1) * W] : n.inE.data();
            int32_t* H = &n.H[c * W]; int32_t* E = &n.E[c * W]; int32_t* F = &n.F[c * W]; int32_t* D = &n.D[c * W];
            H[0] = 0;
            for (size_t j = 1; j < W; j++) {
                if (cells) (*cells)++;
                const int32_t base = (j == 1) ? bonus : std::max(pH[j - 1], 0);
                const int32_t d = base + subst(q[j - 1], n.seq[c]);
                int32_t e = NEG;
                if (pH[j] > 0) e = pH[j] - go;
                if (pE[j] > NEG) e = std::max(e, pE[j] - ge);
                if (e <= 0) e = NEG;
                int32_t f = NEG;
                if (H[j - 1] > 0) f = H[j - 1] - go;
                if (F[j - 1] > NEG) f = std::max(f, F[j - 1] - ge);
                if (f <= 0) f = NEG;
                const int32_t h = std::max(std::max(d, 0), std::max(e, f));
                H[j] = h; E[j] = e; F[j] = f; D[j] = d;
                int32_t cand = h; bool end_diag = false;
                if (j == m && d + bonus >= h) { cand = d + bonus; end_diag = true; }
                if (cand > best) { best = cand; best_u = u; best_c = c; best_j = j; best_end_diag = end_diag; }
            }
        }
    }
    if (best <= 0) return out;
    out.score = best;

    // ---- traceback: diagonal > deletion > insertion, open > extend, stop at a prefix worth <= 0 ----
    struct Step { size_t u; char op; };
    std::vector<Step> steps;                // end -> start
    size_t u = best_u, c = best_c, j = best_j, start_j = 0;
    int state = best_end_diag ? 3 : 0;      // 0 H, 1 E, 2 F, 3 diagonal
    while (true) {
        const NodeDP& n = dp[u];
        const int32_t* pH = c ? &n.H[(c - 1) * W] : n.inH.data();
        if (state == 0) {
            const int32_t h = n.H[c * W + j];
            if (h <= 0) { start_j = j; break; }
            if (n.D[c * W + j] == h) state = 3;
            else if (n.E[c * W + j] == h) state = 1;
            else state = 2;
            continue;
        }
        if (state == 3) {
            steps.push_back({u, 'M'});
            const bool fresh = (j == 1) || !(pH[j - 1] > 0);
            j--;
            if (fresh) { start_j = j; break; }
            if (c > 0) c--; else { u = P.pred[u][n.argH[j]]; c = dp[u].seq.size() - 1; }
            state = 0;
            continue;
        }
        if (state == 1) {
            steps.push_back({u, 'D'});
            const bool open = (pH[j] > 0) && n.E[c * W + j] == pH[j] - go;
            if (c > 0) c--; else { u = P.pred[u][open ? n.argH[j] : n.argE[j]]; c = dp[u].seq.size() - 1; }
            state = open ? 0 : 1;
            continue;
        }
        steps.push_back({u, 'I'});
        const bool open = (n.H[c * W + j - 1] > 0) && n.F[c * W + j] == n.H[c * W + j - 1] - go;
        j--;
        state = open ? 0 : 2;
    }
    std::reverse(steps.begin(), steps.end());

    // ---- steps -> mappings; soft clips as insertions on the first / last mapping ---------------------
    size_t qpos = start_j, i = 0;
    while (i < steps.size()) {
        const size_t node_u = steps[i].u;
        size_t k = i, cols = 0;
        while (k < steps.size() && steps[k].u == node_u) { if (steps[k].op != 'I') cols++; k++; }
        Mapping mp; mp.node = (uint32_t)node_u;
        const size_t end_col = (k == steps.size()) ? best_c + 1 : dp[node_u].seq.size();     // exclusive
        mp.offset = (uint32_t)(end_col - cols);
        if (i == 0 && start_j > 0) { Edit e; e.to_length = (uint32_t)start_j; e.sequence = q.substr(0, start_j); mp.edits.push_back(e); }
        size_t col = mp.offset;
        for (size_t x = i; x < k; x++) {
            const char op = steps[x].op;
            Edit* last = mp.edits.empty() ? nullptr : &mp.edits.back();
            if (op == 'M') {
                if (q[qpos] == dp[node_u].seq[col]) {
                    if (last && last->from_length == last->to_length && last->sequence.empty()) { last->from_length++; last->to_length++; }
                    else { Edit e; e.from_length = e.to_length = 1; mp.edits.push_back(e); }
                } else { Edit e; e.from_length = e.to_length = 1; e.sequence = std::string(1, q[qpos]); mp.edits.push_back(e); }
                qpos++; col++;
            } else if (op == 'I') {
                if (last && last->from_length == 0 && x > i) { last->to_length++; last->sequence.push_back(q[qpos]); }
                else { Edit e; e.to_length = 1; e.sequence = std::string(1, q[qpos]); mp.edits.push_back(e); }
                qpos++;
            } else {
                if (last && last->to_length == 0) last->from_length++;
                else { Edit e; e.from_length = 1; mp.edits.push_back(e); }
                col++;
            }
        }
        out.path.push_back(std::move(mp));
        i = k;
    }
    if (!out.path.empty() && qpos < m) {
        Edit e; e.to_length = (uint32_t)(m - qpos); e.sequence = q.substr(qpos);
        out.path.back().edits.push_back(e);
    }
    return out;
}

} // namespace oracle

// C entry for the stage-level parity tests: one problem.  node[i] oriented nodes in topological
// order; pred / pred_off the CSR of predecessor indices.  Output like oracle_xdrop_pinned, with
// mappings in PROBLEM space (node = index into node[]).
extern "C" int oracle_sw_local(const gb_flat_index* ix, const gb_scores* scores,
                               const uint32_t* node, uint32_t n_nodes, const uint32_t* pred, const uint32_t* pred_off,
                               const uint8_t* query, uint32_t qlen,
                               int32_t* score_out, gb_mapping* mappings, uint32_t mapping_cap, uint32_t* n_mappings,
                               uint32_t* edits, uint32_t edit_cap, uint32_t* n_edits, uint64_t* cells_out) {
    oracle::Graph g(ix);
    oracle::DagProblem P;
    P.node.assign(node, node + n_nodes); P.pred.resize(n_nodes);
    for (uint32_t u = 0; u < n_nodes; u++) P.pred[u].assign(pred + pred_off[u], pred + pred_off[u + 1]);
    uint64_t cells = 0;
    oracle::LocalAlignmentResult a = oracle::sw_local_dag(g, *scores, P, std::string((const char*)query, qlen), &cells);
    if (cells_out) *cells_out = cells;
    *score_out = a.score;
    if (a.path.size() > mapping_cap) return -1;
    uint32_t ne = 0;
    for (size_t i = 0; i < a.path.size(); i++) {
        mappings[i].node = a.path[i].node; mappings[i].offset = (uint16_t)a.path[i].offset; mappings[i].n_edits = (uint16_t)a.path[i].edits.size();
        for (const oracle::Edit& e : a.path[i].edits) {
            if (ne >= edit_cap) return -1;
            uint32_t word;
            if (e.from_length == e.to_length && e.sequence.empty()) word = (e.from_length << 4) | GB_EDIT_MATCH;
            else if (e.from_length == e.to_length) {
                uint32_t b = 0; switch (e.sequence[0]) { case 'C': b = 1; break; case 'G': b = 2; break; case 'T': b = 3; break; default: b = 0; }
                word = (e.from_length << 4) | (b << 2) | GB_EDIT_SUB;
            } else if (e.from_length == 0) word = (e.to_length << 4) | GB_EDIT_INS;
            else word = (e.from_length << 4) | GB_EDIT_DEL;
            edits[ne++] = word;
        }
    }
    *n_mappings = (uint32_t)a.path.size(); *n_edits = ne;
    return 0;
}
