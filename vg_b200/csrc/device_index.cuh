// device_index.cuh — HBM-resident flat index and warp-cooperative GBWT primitives.
//
// Everything here is warp-synchronous: all 32 lanes of a warp call each function with
// warp-uniform arguments and receive warp-uniform results.  A GBWT record is decoded by
// the whole warp (lane j <-> run j, lane e <-> edge e) so LF-mapping and the
// bidirectional range update cost one coalesced load of the record blob plus a few
// shuffles instead of a serial run scan.
//
// Semantics restate jltsiren/gbwt @ c2e0199 (absent from /root/reference):
//   bdExtendForward / follow_paths as used at vg gbwt_extender.cpp:608, :652.
#pragma once
#include "giraffe_b200.h"
#include <cuda_runtime.h>
#include <cstdint>

namespace gb {

constexpr unsigned FULL = 0xffffffffu;

struct DevIndex {
    const gb_node_rec* nodes;
    const uint8_t* seq;
    const uint32_t* gbwt;
    const gb_dist_payload* dist;
    const gb_min_cell* table;
    const gb_hit* hits;
    uint64_t table_mask;
    uint32_t n_nodes, k, w;
    const uint32_t* slot_order;      // node ids sorted by (component, slot, place in the site): a topological order of each chain
    uint32_t n_ids;
    const gb_slot_rec* slots; uint32_t n_slots;      // site tables (gb_dist_payload); n_slots == 0: none
    const uint16_t* site_dist;
};

// Minimum distance from the END of node id_u to the START of node id_v when both lie in the same slot (forward strand);
// -1: not reachable.  pu / pv are their payloads as uint4 (x_in, x_out, slot, allele | component << 16).
__device__ __forceinline__ int64_t site_distance(const DevIndex& ix, const uint4& pu, const uint4& pv) {
    const uint32_t slot = pu.z;
    if (slot >= ix.n_slots) return -1;
    const gb_slot_rec sr = ix.slots[slot];
    const uint32_t a = pu.w & 0xFFFFu, b = pv.w & 0xFFFFu;
    if (sr.table_off == 0xFFFFFFFFu || a >= sr.n || b >= sr.n) return -1;
    const uint16_t d = __ldg(ix.site_dist + sr.table_off + (size_t)a * sr.n + b);
    return d == 0xFFFFu ? -1 : (int64_t)d;
}

struct DevScores { int match, mismatch, gap_open, gap_extend, full_length_bonus; };

// gbwt::BidirectionalState with closed int32 ranges (empty: lo > hi).
struct BdState {
    uint32_t fnode; int32_t flo, fhi;
    uint32_t bnode; int32_t blo, bhi;
    __device__ __forceinline__ int32_t size() const { return fhi - flo + 1; }
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ gb_node_rec load_node(const DevIndex& ix, uint32_t v) {
    // one 128-bit load of the 16-byte node record
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(ix.nodes) + v);
    gb_node_rec n; n.seq_off = r.x; n.rec_off = r.y; n.len = r.z; n.size = r.w;
    return n;
}

// minimum_distance(pos1, pos2) on cut-style positions through the distance payload; unreachable
// is size_t max stored into int64_t (= -1), as in distance_between (:3879-3884).
__device__ inline int64_t oriented_distance(const DevIndex& ix, uint32_t node_a, uint32_t off_a, uint32_t node_b, uint32_t off_b) {
    const int64_t UNREACHABLE = -1;
    if ((node_a & 1u) != (node_b & 1u)) return UNREACHABLE;
    uint32_t src = node_a, dst = node_b; int64_t src_off = off_a, dst_off = off_b;
    if (node_a & 1u) {
        src = node_b; dst = node_a;
        src_off = (int64_t)load_node(ix, node_b).len - (int64_t)off_b;
        dst_off = (int64_t)load_node(ix, node_a).len - (int64_t)off_a;
    }
    const int64_t src_len = load_node(ix, src).len;
    const uint4 ps = __ldg(reinterpret_cast<const uint4*>(ix.dist) + (src >> 1));
    const uint4 pd = __ldg(reinterpret_cast<const uint4*>(ix.dist) + (dst >> 1));
    if ((ps.w >> 16) != (pd.w >> 16)) return UNREACHABLE;               // component
    if ((src >> 1) == (dst >> 1)) return dst_off >= src_off ? dst_off - src_off : UNREACHABLE;
    if (ps.z < pd.z) return (src_len - src_off) + ((int64_t)(int32_t)pd.x - (int64_t)(int32_t)ps.y) + dst_off;
    if (ps.z == pd.z) { const int64_t t = site_distance(ix, ps, pd); if (t >= 0) return (src_len - src_off) + t + dst_off; }
    return UNREACHABLE;
}


__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

__device__ __forceinline__ int warp_incl_scan(int v) {
    const int lane = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(FULL, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// Decoded view of (up to 32 edges of) a GBWT record for a visit range [lo, hi]:
// lane e holds edge e: successor `to`, first position `first` of the extended range in the
// successor's record, number `cnt` of visits in [lo,hi] taking the edge, and `rev` = the
// bidirectional reverse offset (visits whose successor sorts before `to` on the reverse
// strand).  n_edges is warp-uniform.
struct EdgeFan {
    uint32_t to; int32_t first; int32_t cnt; int32_t rev;
    uint32_t n_edges;
};

// Generic (any outdegree / any run count) single-edge query, used when a record has more
// than 32 edges.  Warp-uniform result.
__device__ inline void record_edge_query(const uint32_t* rec, uint32_t n_edges, uint32_t n_runs,
                                         int32_t lo, int32_t hi, uint32_t e,
                                         int32_t& below, int32_t& cnt) {
    const uint32_t* runs = rec + 2 + 2 * n_edges;
    const int lane = lane_id();
    int32_t base = 0; below = 0; cnt = 0;
    for (uint32_t rb = 0; rb < n_runs; rb += 32) {
        uint32_t w = (rb + lane < n_runs) ? __ldg(runs + rb + lane) : 0u;
        int32_t len = (int32_t)(w >> 10); uint32_t rr = w & 1023u;
        int32_t end = base + warp_incl_scan(len);
        int32_t start = end - len;
        int32_t b = min(max(lo - start, 0), len);
        int32_t o = max(min(end, hi + 1) - max(start, lo), 0);
        bool mine = (rb + lane < n_runs) && rr == e;
        below += warp_sum(mine ? b : 0);
        cnt += warp_sum(mine ? o : 0);
        base = __shfl_sync(FULL, end, 31);
        if (base > hi) break;
    }
}

// Decode the record of node v for the range [lo, hi] (edges eb .. eb+31).
__device__ inline EdgeFan record_fan(const DevIndex& ix, const gb_node_rec& nr, int32_t lo, int32_t hi) {
    EdgeFan f; f.to = 0; f.first = 0; f.cnt = 0; f.rev = 0; f.n_edges = 0;
    if (nr.size == 0) return f;
    const uint32_t* rec = ix.gbwt + nr.rec_off;
    const int lane = lane_id();
    uint32_t head = lane < 2 ? __ldg(rec + lane) : 0u;
    const uint32_t n_edges = __shfl_sync(FULL, head, 0);
    const uint32_t n_runs = __shfl_sync(FULL, head, 1);
    f.n_edges = n_edges;
    if (n_edges == 1 && n_runs == 1) {
        // the usual record away from variant sites: one successor, one run.  Every visit of [lo, hi] takes the edge, so the
        // extended range is the edge offset plus the visits before lo — no scan over runs, no vote over edges.
        const uint2 e = __ldg(reinterpret_cast<const uint2*>(rec + 2));
        const int32_t len = (int32_t)(__ldg(rec + 4) >> 10);
        const int32_t b = min(max(lo, 0), len);
        const int32_t o = max(min(len, hi + 1) - max(0, lo), 0);
        f.to = e.x; f.first = (int32_t)e.y + b; f.cnt = o; f.rev = 0;
        return f;
    }
    if (n_edges <= 32) {
        uint32_t my_to = 0, my_off = 0;
        if (lane < n_edges) {
            const uint2 e = __ldg(reinterpret_cast<const uint2*>(rec + 2) + lane);
            my_to = e.x; my_off = e.y;
        }
        const uint32_t* runs = rec + 2 + 2 * n_edges;
        int32_t my_below = 0, my_cnt = 0, base = 0;
        for (uint32_t rb = 0; rb < n_runs; rb += 32) {
            uint32_t w = (rb + lane < n_runs) ? __ldg(runs + rb + lane) : 0u;
            int32_t len = (int32_t)(w >> 10); uint32_t rr = w & 1023u;
            int32_t end = base + warp_incl_scan(len);
            int32_t start = end - len;
            int32_t b = min(max(lo - start, 0), len);
            int32_t o = max(min(end, hi + 1) - max(start, lo), 0);
            const uint32_t in_chunk = min(32u, n_runs - rb);
            for (uint32_t t = 0; t < in_chunk; t++) {
                uint32_t rt = __shfl_sync(FULL, rr, t);
                int32_t bt = __shfl_sync(FULL, b, t);
                int32_t ot = __shfl_sync(FULL, o, t);
                if (rt == (uint32_t)lane) { my_below += bt; my_cnt += ot; }
            }
            base = __shfl_sync(FULL, end, 31);
            if (base > hi) break;   // runs past the range contribute nothing
        }
        int32_t rev = 0;
        for (uint32_t t = 0; t < n_edges; t++) {
            uint32_t tt = __shfl_sync(FULL, my_to, t);
            int32_t ct = __shfl_sync(FULL, my_cnt, t);
            if ((tt ^ 1u) < (my_to ^ 1u)) rev += ct;
        }
        f.to = my_to; f.first = (int32_t)my_off + my_below; f.cnt = my_cnt; f.rev = rev;
    }
    return f;
}

// Slow generic path for records with more than 32 edges: edge e of the record, uniform.
__device__ inline void record_edge_generic(const DevIndex& ix, const gb_node_rec& nr, int32_t lo, int32_t hi,
                                           uint32_t e, uint32_t& to, int32_t& first, int32_t& cnt, int32_t& rev) {
    const uint32_t* rec = ix.gbwt + nr.rec_off;
    const uint32_t n_edges = __ldg(rec), n_runs = __ldg(rec + 1);
    to = __ldg(rec + 2 + 2 * e);
    uint32_t off = __ldg(rec + 3 + 2 * e);
    int32_t below;
    record_edge_query(rec, n_edges, n_runs, lo, hi, e, below, cnt);
    first = (int32_t)off + below;
    rev = 0;
    if (cnt > 0 && to != 0) {
        for (uint32_t e2 = 0; e2 < n_edges; e2++) {
            uint32_t to2 = __ldg(rec + 2 + 2 * e2);
            if ((to2 ^ 1u) < (to ^ 1u)) {
                int32_t b2, c2;
                record_edge_query(rec, n_edges, n_runs, lo, hi, e2, b2, c2);
                rev += c2;
            }
        }
    }
}

__device__ __forceinline__ BdState bd_state_of(const gb_node_rec& nr, uint32_t v) {
    BdState s;
    s.fnode = v; s.flo = 0; s.fhi = (int32_t)nr.size - 1;
    s.bnode = v ^ 1u; s.blo = 0; s.bhi = (int32_t)nr.size - 1;
    return s;
}

__device__ __forceinline__ BdState bd_flip(const BdState& s) {
    BdState r; r.fnode = s.bnode; r.flo = s.blo; r.fhi = s.bhi; r.bnode = s.fnode; r.blo = s.flo; r.bhi = s.fhi;
    return r;
}

// Apply edge (to, first, cnt, rev) to a state (already flipped for backward extension).
__device__ __forceinline__ BdState bd_apply(const BdState& s, uint32_t to, int32_t first, int32_t cnt, int32_t rev) {
    BdState n;
    n.fnode = to; n.flo = first; n.fhi = first + cnt - 1;
    n.bnode = s.bnode; n.blo = s.blo + rev; n.bhi = n.blo + cnt - 1;
    return n;
}

} // namespace gb
