// extend.cuh — warp-synchronous haplotype-consistent gapless extension (device side).
//
// One warp owns one work item = one call of vg::GaplessExtender::extend
// (gbwt_extender.cpp:533-737): a read plus the seeds of one cluster.  The warp
//   * stages the masked read in shared memory (ReadMasker, gbwt_extender.cpp:160-170),
//   * runs the reference's best-first search per seed with the frontier (priority queue)
//     in a per-warp HBM workspace and the partial paths in a parent-pointer arena,
//   * compares 32 read/graph bases per step (one <=32-bp node per step) with a ballot and
//     resolves the mismatch budget with popc / fns instead of a byte loop,
//   * decodes GBWT records cooperatively (device_index.cuh),
//   * and finishes with the reference's post-processing (full-length selection with the
//     overlap filter, or duplicate removal + mismatch trimming) on the output records.
//
// Integer only.  Results are bit-identical to the oracle (oracle/extend.cpp).
#pragma once
#include "device_index.cuh"

namespace gb {

constexpr uint32_t NONE = 0xffffffffu;
constexpr uint32_t F_LEFT_FULL = 1u, F_RIGHT_FULL = 2u, F_LEFT_MAX = 4u, F_RIGHT_MAX = 8u;

// One frontier entry (64 B): a GaplessExtension under construction.
struct __align__(16) QEntry {
    uint32_t fnode; int32_t flo, fhi;
    uint32_t bnode; int32_t blo, bhi;
    uint32_t read_lo, read_hi;
    int32_t  score;
    uint32_t number;           // extension_number: queue tie-break (gbwt_extender.cpp:571)
    uint32_t internal_score;   // mismatches so far
    uint32_t old_score;        // mismatches before the current flank
    uint32_t offset;           // offset in the first node of the path
    uint32_t flags;
    uint32_t right_tail;       // arena index of the last right-extension node (NONE: seed node)
    uint32_t left_head;        // arena index of the first left-extension node  (NONE: seed node)
};
static_assert(sizeof(QEntry) == 64, "QEntry must be 64 bytes");

struct ArenaNode { uint32_t node, parent; };

struct ExtendWorkspace {
    QEntry* queue;        // [n_warps * q_cap]
    ArenaNode* arena;     // [n_warps * a_cap]
    uint32_t q_cap, a_cap;
};

// large-stride outputs for the few items (clusters of repeats) that overflow the regular per-item strides
struct ExtendBig {
    uint32_t* list; uint32_t* count; uint32_t cap;          // items to redo (device list + counter)
    uint32_t* big_of;                                       // [n_items] slot of an item in the large pools, 0xffffffff: regular
    gb_extension* ext; uint32_t* path; uint32_t* mism;
    uint32_t max_ext, path_cap, mism_cap;
};

struct ExtendParams {
    DevScores sc;
    uint32_t max_mismatches;
    float overlap_threshold_unused;
    double overlap_threshold;
    uint32_t trim;
    uint32_t max_ext, path_cap, mism_cap;
};

__device__ __forceinline__ void set_score(QEntry& e, const DevScores& sc) {
    // gbwt_extender.cpp:201-209
    int32_t s = (int32_t)(e.read_hi - e.read_lo) * sc.match;
    s -= (int32_t)e.internal_score * (sc.match + sc.mismatch);
    s += (e.flags & F_LEFT_FULL) ? sc.full_length_bonus : 0;
    s += (e.flags & F_RIGHT_FULL) ? sc.full_length_bonus : 0;
    e.score = s;
}

__device__ __forceinline__ uint32_t mismatch_limit_of(uint32_t max_mismatches, uint32_t old_score) {
    // gbwt_extender.cpp:605-607
    return max(max_mismatches + 1u, max_mismatches / 2u + old_score + 1u);
}

// Compare read[rpos + i] with node[npos + i], i in [0, count), forward.  Stops before the
// mismatch that would make internal >= limit (limit == 0: unlimited).  Returns the number
// of bases consumed; updates `internal`.  Warp-uniform.
__device__ __forceinline__ uint32_t match_fwd(const uint8_t* sread, uint32_t rpos, const uint8_t* __restrict__ nseq,
                                              uint32_t npos, uint32_t count, uint32_t& internal, uint32_t limit) {
    const int lane = lane_id();
    uint32_t consumed = 0;
    while (consumed < count) {
        const uint32_t i = consumed + lane;
        const bool active = i < count;
        uint8_t r = 0, t = 0;
        if (active) { r = sread[rpos + i]; t = __ldg(nseq + npos + i); }
        const uint32_t m = __ballot_sync(FULL, active && r != t);
        const uint32_t chunk = min(32u, count - consumed);
        const uint32_t c = __popc(m);
        if (limit == 0) { internal += c; consumed += chunk; continue; }
        const uint32_t budget = (internal + 1 >= limit) ? 0u : (limit - 1u - internal);
        if (c <= budget) { internal += c; consumed += chunk; }
        else {
            const uint32_t stop = __fns(m, 0, (int)budget + 1);
            internal += budget; consumed += stop;
            break;
        }
    }
    return consumed;
}

// Compare read[rpos - 1 - i] with node[npos - 1 - i], i in [0, count), backward.
__device__ __forceinline__ uint32_t match_bwd(const uint8_t* sread, uint32_t rpos, const uint8_t* __restrict__ nseq,
                                              uint32_t npos, uint32_t count, uint32_t& internal, uint32_t limit) {
    const int lane = lane_id();
    uint32_t consumed = 0;
    while (consumed < count) {
        const uint32_t i = consumed + lane;
        const bool active = i < count;
        uint8_t r = 0, t = 0;
        if (active) { r = sread[rpos - 1 - i]; t = __ldg(nseq + npos - 1 - i); }
        const uint32_t m = __ballot_sync(FULL, active && r != t);
        const uint32_t chunk = min(32u, count - consumed);
        const uint32_t c = __popc(m);
        const uint32_t budget = (internal + 1 >= limit) ? 0u : (limit - 1u - internal);
        if (c <= budget) { internal += c; consumed += chunk; }
        else {
            const uint32_t stop = __fns(m, 0, (int)budget + 1);
            internal += budget; consumed += stop;
            break;
        }
    }
    return consumed;
}

// ---- frontier (priority queue) -------------------------------------------------------
// Pop order of std::priority_queue<std::pair<GaplessExtension,size_t>> is fully determined
// by the key (score, number) because numbers are unique, so any exact arg-max reproduces it.

__device__ __forceinline__ void q_store(QEntry* slot, const QEntry& e) {
    // warp-uniform entry: lane 0 writes the four 16-byte quarters
    if (lane_id() == 0) {
        uint4* dst = reinterpret_cast<uint4*>(slot);
        dst[0] = make_uint4(e.fnode, (uint32_t)e.flo, (uint32_t)e.fhi, e.bnode);
        dst[1] = make_uint4((uint32_t)e.blo, (uint32_t)e.bhi, e.read_lo, e.read_hi);
        dst[2] = make_uint4((uint32_t)e.score, e.number, e.internal_score, e.old_score);
        dst[3] = make_uint4(e.offset, e.flags, e.right_tail, e.left_head);
    }
}

__device__ __forceinline__ QEntry q_load(const QEntry* slot) {
    QEntry e;
    uint4* dst = reinterpret_cast<uint4*>(&e);
    const uint4* src = reinterpret_cast<const uint4*>(slot);
#pragma unroll
    for (int i = 0; i < 4; i++) dst[i] = src[i];
    return e;
}

__device__ inline QEntry q_pop(QEntry* queue, uint32_t& qn) {
    const int lane = lane_id();
    if (qn == 1) { QEntry only = q_load(queue); qn = 0; __syncwarp(); return only; }      // a linear walk keeps one entry
    long long best = LLONG_MIN; uint32_t best_idx = 0;
    for (uint32_t j = lane; j < qn; j += 32) {
        const long long key = (long long)queue[j].score * 4294967296LL + (long long)queue[j].number;
        if (key > best) { best = key; best_idx = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long ob = __shfl_xor_sync(FULL, best, o);
        const uint32_t oi = __shfl_xor_sync(FULL, best_idx, o);
        if (ob > best) { best = ob; best_idx = oi; }
    }
    __syncwarp();
    QEntry e = q_load(queue + best_idx);
    qn--;
    if (best_idx != qn) {
        QEntry last = q_load(queue + qn);
        __syncwarp();
        q_store(queue + best_idx, last);
    }
    __syncwarp();
    return e;
}

// ---- helpers on finished extensions (output records) ---------------------------------

// GaplessExtension::contains, gbwt_extender.cpp:41-53.  Lane-0 style serial walk, uniform.
__device__ inline bool ext_contains(const DevIndex& ix, const gb_extension& e, const uint32_t* path_pool,
                                    uint32_t node, int32_t diag) {
    uint32_t read_offset = e.read_lo, node_offset = e.offset;
    for (uint32_t i = 0; i < e.path_len; i++) {
        const uint32_t h = path_pool[e.path_off + i];
        const uint32_t nlen = load_node(ix, h).len;
        const uint32_t len = min(nlen - node_offset, e.read_hi - read_offset);
        if (h == node && (int32_t)read_offset - (int32_t)node_offset == diag) return true;
        read_offset += len; node_offset = 0;
    }
    return false;
}

// GaplessExtension::overlap, gbwt_extender.cpp:89-117.
__device__ inline uint32_t ext_overlap(const DevIndex& ix, const gb_extension& a, const gb_extension& b,
                                       const uint32_t* path_pool) {
    uint32_t result = 0;
    uint32_t a_pos = a.read_lo, b_pos = b.read_lo;
    uint32_t ai = 0, bi = 0;
    uint32_t a_off = a.offset, b_off = b.offset;
    while (a_pos < a.read_hi && b_pos < b.read_hi) {
        const uint32_t an = path_pool[a.path_off + ai], bn = path_pool[b.path_off + bi];
        if (a_pos == b_pos && an == bn && a_off == b_off) {
            const uint32_t nlen = load_node(ix, an).len;
            const uint32_t len = min(min(nlen - a_off, a.read_hi - a_pos), b.read_hi - b_pos);
            result += len; a_pos += len; b_pos += len; ai++; bi++; a_off = 0; b_off = 0;
        } else if (a_pos <= b_pos) {
            a_pos += load_node(ix, an).len - a_off; ai++; a_off = 0;
        } else {
            b_pos += load_node(ix, bn).len - b_off; bi++; b_off = 0;
        }
    }
    return result;
}

// remove_duplicates ordering, gbwt_extender.cpp:333-350.
__device__ __forceinline__ bool dup_less(const gb_extension& a, const gb_extension& b) {
    if (a.read_lo != b.read_lo) return a.read_lo < b.read_lo;
    if (a.read_hi != b.read_hi) return a.read_hi < b.read_hi;
    if (a.bwd_node != b.bwd_node) return a.bwd_node < b.bwd_node;
    if (a.fwd_node != b.fwd_node) return a.fwd_node < b.fwd_node;
    if (a.bwd_lo != b.bwd_lo) return (int32_t)a.bwd_lo < (int32_t)b.bwd_lo;
    if (a.bwd_hi != b.bwd_hi) return (int32_t)a.bwd_hi < (int32_t)b.bwd_hi;
    if (a.fwd_lo != b.fwd_lo) return (int32_t)a.fwd_lo < (int32_t)b.fwd_lo;
    if (a.fwd_hi != b.fwd_hi) return (int32_t)a.fwd_hi < (int32_t)b.fwd_hi;
    return a.offset < b.offset;
}
__device__ __forceinline__ bool ext_equal(const gb_extension& a, const gb_extension& b) {
    return a.read_lo == b.read_lo && a.read_hi == b.read_hi && a.offset == b.offset &&
           a.fwd_node == b.fwd_node && a.fwd_lo == b.fwd_lo && a.fwd_hi == b.fwd_hi &&
           a.bwd_node == b.bwd_node && a.bwd_lo == b.bwd_lo && a.bwd_hi == b.bwd_hi;
}

// Stable insertion sort + dedupe of the output records by lane 0 (n is small).
static __device__ __noinline__ uint32_t remove_duplicates(gb_extension* ext, uint32_t n) {
    if (lane_id() == 0) {
        for (uint32_t i = 1; i < n; i++) {
            gb_extension key = ext[i];
            uint32_t j = i;
            while (j > 0 && dup_less(key, ext[j - 1])) { ext[j] = ext[j - 1]; j--; }
            ext[j] = key;
        }
        uint32_t tail = 0;
        for (uint32_t i = 0; i < n; i++) {
            if (ext[i].read_hi == ext[i].read_lo) continue;
            if (tail == 0 || !ext_equal(ext[i], ext[tail - 1])) {
                if (i > tail) ext[tail] = ext[i];
                tail++;
            }
        }
        n = tail;
    }
    n = __shfl_sync(FULL, n, 0);
    __syncwarp();
    return n;
}

// find_mismatches for one extension (gbwt_extender.cpp:368-387), warp-parallel.
// Returns false when the mismatch pool overflows.
__device__ inline bool find_mismatches(const DevIndex& ix, gb_extension& e, const uint8_t* sread,
                                       const uint32_t* path_pool, uint32_t* mism_pool,
                                       uint32_t& mism_used, uint32_t mism_end) {
    const int lane = lane_id();
    e.mism_off = mism_used; e.mism_len = 0;
    if (e.mismatches == 0) return true;
    uint32_t node_offset = e.offset, read_offset = e.read_lo;
    for (uint32_t i = 0; i < e.path_len && read_offset < e.read_hi; i++) {
        const gb_node_rec nr = load_node(ix, path_pool[e.path_off + i]);
        const uint32_t count = min(nr.len - node_offset, e.read_hi - read_offset);
        for (uint32_t c0 = 0; c0 < count; c0 += 32) {
            const uint32_t j = c0 + lane;
            bool mm = false;
            if (j < count) mm = sread[read_offset + j] != __ldg(ix.seq + nr.seq_off + node_offset + j);
            const uint32_t m = __ballot_sync(FULL, mm);
            const uint32_t c = __popc(m);
            if (mism_used + c > mism_end) return false;
            if (mm) mism_pool[mism_used + __popc(m & ((1u << lane) - 1u))] = read_offset + j;
            mism_used += c;
        }
        read_offset += count; node_offset = 0;
    }
    e.mism_len = mism_used - e.mism_off;
    __syncwarp();
    return true;
}

// bd_find(path), gbwt_extender.cpp:514.  Warp-cooperative walk.
__device__ inline BdState bd_find(const DevIndex& ix, const uint32_t* path, uint32_t n) {
    gb_node_rec nr = load_node(ix, path[0]);
    BdState s = bd_state_of(nr, path[0]);
    for (uint32_t i = 1; i < n; i++) {
        const uint32_t target = path[i];
        EdgeFan f = record_fan(ix, nr, s.flo, s.fhi);
        uint32_t to = 0; int32_t first = 0, cnt = 0, rev = 0;
        if (f.n_edges <= 32) {
            const uint32_t hit = __ballot_sync(FULL, (uint32_t)lane_id() < f.n_edges && f.to == target);
            if (hit) {
                const int src = __ffs(hit) - 1;
                to = target; first = __shfl_sync(FULL, f.first, src); cnt = __shfl_sync(FULL, f.cnt, src); rev = __shfl_sync(FULL, f.rev, src);
            }
        } else {
            for (uint32_t e = 0; e < f.n_edges; e++) {
                uint32_t t2; int32_t f2, c2, r2;
                record_edge_generic(ix, nr, s.flo, s.fhi, e, t2, f2, c2, r2);
                if (t2 == target) { to = t2; first = f2; cnt = c2; rev = r2; break; }
            }
        }
        if (to != target || cnt <= 0) { s.flo = 0; s.fhi = -1; s.blo = 0; s.bhi = -1; s.fnode = target; return s; }
        s = bd_apply(s, to, first, cnt, rev);
        nr = load_node(ix, target);
    }
    return s;
}

// trim_mismatches (gbwt_extender.cpp:421-529).  Serial arithmetic is warp-uniform (every
// lane computes the same values from the same memory); bd_find is cooperative.
__device__ inline bool trim_mismatches(const DevIndex& ix, gb_extension& e, const DevScores& sc,
                                       const uint32_t* path_pool, const uint32_t* mism_pool) {
    if (e.mism_len == 0) return false;
    const uint32_t* mm = mism_pool + e.mism_off;
    uint32_t mi = 0;
    uint32_t cur_lo = e.read_lo, cur_hi = mm[0];
    int32_t cur_score = (int32_t)(cur_hi - cur_lo) * sc.match;
    if (e.flags & GB_EXT_LEFT_FULL) cur_score += sc.full_length_bonus;
    uint32_t best_lo = cur_lo, best_hi = cur_hi; int32_t best_score = cur_score;
    while (mi < e.mism_len) {
        if (cur_score >= sc.mismatch) { cur_hi++; cur_score -= sc.mismatch; }
        else { cur_lo = cur_hi = mm[mi] + 1; cur_score = 0; }
        mi++;
        if (mi == e.mism_len) {
            const uint32_t length = e.read_hi - cur_hi;
            cur_hi = e.read_hi; cur_score += (int32_t)length * sc.match;
            if (e.flags & GB_EXT_RIGHT_FULL) cur_score += sc.full_length_bonus;
        } else {
            const uint32_t length = mm[mi] - cur_hi;
            cur_hi = mm[mi]; cur_score += (int32_t)length * sc.match;
        }
        if (cur_score > best_score || (cur_score > 0 && cur_score == best_score && (cur_hi - cur_lo) > (best_hi - best_lo))) {
            best_lo = cur_lo; best_hi = cur_hi; best_score = cur_score;
        }
    }
    if (best_lo == e.read_lo && best_hi == e.read_hi) return false;
    if (best_hi == best_lo) {
        e.path_len = 0; e.read_lo = best_lo; e.read_hi = best_hi; e.mism_len = 0; e.score = 0; e.flags = 0;
        return true;
    }
    if (best_lo > e.read_lo) e.flags &= ~GB_EXT_LEFT_FULL;
    if (best_hi < e.read_hi) e.flags &= ~GB_EXT_RIGHT_FULL;
    uint32_t node_offset = e.offset, read_offset = e.read_lo;
    e.read_lo = best_lo; e.read_hi = best_hi; e.score = best_score;
    uint32_t head = 0;
    while (head < e.path_len) {
        const uint32_t node_length = load_node(ix, path_pool[e.path_off + head]).len;
        read_offset += node_length - node_offset;
        node_offset = 0;
        if (read_offset > e.read_lo) { e.offset = node_length - (read_offset - e.read_lo); break; }
        head++;
    }
    uint32_t tail = head + 1;
    while (read_offset < e.read_hi) { read_offset += load_node(ix, path_pool[e.path_off + tail]).len; tail++; }
    if (head > 0 || tail < e.path_len) {
        e.path_off += head; e.path_len = tail - head;
        const BdState s = bd_find(ix, path_pool + e.path_off, e.path_len);
        e.fwd_node = s.fnode; e.fwd_lo = (uint32_t)s.flo; e.fwd_hi = (uint32_t)s.fhi;
        e.bwd_node = s.bnode; e.bwd_lo = (uint32_t)s.blo; e.bwd_hi = (uint32_t)s.bhi;
    }
    uint32_t mh = 0;
    while (mh < e.mism_len && mm[mh] < e.read_lo) mh++;
    uint32_t mt = mh;
    while (mt < e.mism_len && mm[mt] < e.read_hi) mt++;
    e.mism_off += mh; e.mism_len = mt - mh;
    return true;
}

// --------------------------------------------------------------------------------------
// GaplessExtender::extend for one work item, executed by one warp, in two parts so that the kernels of the mapping
// pipeline keep the search loop and the post-processing in separate (instruction-cache sized) kernels:
//   extend_search  gbwt_extender.cpp:540-700  best-first search per seed; writes one raw record per seed that produced
//                  an extension (path in path_pool, `mismatches` = mismatch count of the search, no mismatch list yet)
//   extend_finish  gbwt_extender.cpp:702-736  full-length selection with the overlap filter, or duplicate removal +
//                  mismatch lists + trimming, on those records
// extend_item runs both (rescue, which extends inside the align kernel).
// `ext`, `path_pool`, `mism_pool` point at this item's output regions (capacities in p).
// Returns the number of extensions; *status receives a GB_ITEM_* code.
// --------------------------------------------------------------------------------------
__device__ inline uint32_t extend_search(const DevIndex& ix, const ExtendParams& p,
                                         const uint8_t* sread, uint32_t read_len,
                                         const gb_seed* seeds, uint32_t n_seeds,
                                         QEntry* queue, uint32_t q_cap, ArenaNode* arena, uint32_t a_cap,
                                         gb_extension* ext, uint32_t* path_pool,
                                         uint32_t* status_out) {
    const int lane = lane_id();
    uint32_t status = GB_ITEM_OK;
    uint32_t n_res = 0, path_used = 0;
    uint32_t best_alignment = NONE;          // index into ext[]
    uint32_t best_alignment_mm = 0;
    if (n_seeds == 0 || read_len == 0) { *status_out = status; return 0; }

    // canonical seed order: ascending (node, diag), duplicates collapsed
    unsigned long long last_key = 0; bool have_last = false;
    // bit c of `dead`: my seed of chunk c (seed lane + 32 c) lies on the exact full-length best alignment
    // (gbwt_extender.cpp:553-557), decided for all seeds at once when that alignment appears
    uint32_t dead = 0;
    const bool dead_ok = n_seeds <= 1024;
    while (true) {
        unsigned long long mine = ~0ull;
        for (uint32_t j = lane, c = 0; j < n_seeds; j += 32, c++) {
            if (dead_ok && ((dead >> c) & 1u)) continue;
            const gb_seed s = seeds[j];
            const unsigned long long key = ((unsigned long long)s.node << 32) | (uint32_t)(s.diag ^ 0x80000000);
            if ((!have_last || key > last_key) && key < mine) mine = key;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(FULL, mine, o);
            mine = other < mine ? other : mine;
        }
        if (mine == ~0ull) break;
        last_key = mine; have_last = true;
        const uint32_t seed_node = (uint32_t)(mine >> 32);
        const int32_t seed_diag = (int32_t)((uint32_t)mine ^ 0x80000000u);
        if (seed_node < 2 || seed_node >= ix.n_nodes) continue;

        // gbwt_extender.cpp:553-557: skip seeds contained in an exact full-length alignment
        if (!dead_ok && best_alignment != NONE && best_alignment_mm == 0) {
            if (ext_contains(ix, ext[best_alignment], path_pool, seed_node, seed_diag)) continue;
        }

        QEntry best; best.score = INT_MIN; best.read_lo = best.read_hi = 0; best.flags = 0;
        best.internal_score = 0xffffffffu; best.old_score = 0xffffffffu;
        best.left_head = best.right_tail = NONE; best.offset = 0; best.number = 0;
        best.fnode = best.bnode = 0; best.flo = best.blo = 0; best.fhi = best.bhi = -1;
        bool have_best = false;

        uint32_t qn = 0, an = 0, number = 0;
        const gb_node_rec seed_rec = load_node(ix, seed_node);
        if (seed_rec.len == 0) continue;
        {
            // gbwt_extender.cpp:573-593
            const uint32_t read_offset = seed_diag < 0 ? 0u : (uint32_t)seed_diag;
            const uint32_t node_offset = seed_diag < 0 ? (uint32_t)(-seed_diag) : 0u;
            if (read_offset > read_len || node_offset > seed_rec.len) continue;
            QEntry m;
            const BdState s0 = bd_state_of(seed_rec, seed_node);
            m.fnode = s0.fnode; m.flo = s0.flo; m.fhi = s0.fhi; m.bnode = s0.bnode; m.blo = s0.blo; m.bhi = s0.bhi;
            m.read_lo = read_offset; m.read_hi = read_offset; m.offset = node_offset;
            m.internal_score = 0; m.old_score = 0; m.flags = 0; m.number = number++;
            m.left_head = NONE; m.right_tail = NONE;
            const uint32_t count = min(read_len - read_offset, seed_rec.len - node_offset);
            uint32_t internal = 0;
            m.read_hi += match_fwd(sread, read_offset, ix.seq + seed_rec.seq_off, node_offset, count, internal, 0);
            m.internal_score = internal; m.old_score = internal;
            if (m.read_lo == 0) m.flags |= F_LEFT_FULL | F_LEFT_MAX;
            if (m.read_hi >= read_len) m.flags |= F_RIGHT_FULL | F_RIGHT_MAX;
            set_score(m, p.sc);
            q_store(queue + qn, m); qn++;
            __syncwarp();
        }

        while (qn > 0) {
            QEntry curr = q_pop(queue, qn);

            // Cases 1 and 2 (gbwt_extender.cpp:602-643 right, :646-686 left) share one body: the kernel is
            // instruction-fetch bound, so the walk over the GBWT edges exists once and the direction is data.
            const bool go_right = !(curr.flags & F_RIGHT_MAX);
            if (go_right || !(curr.flags & F_LEFT_MAX)) {
                uint32_t num_extensions = 0; bool found_extension = false;
                const uint32_t limit = mismatch_limit_of(p.max_mismatches, curr.old_score);
                BdState cs; cs.fnode = curr.fnode; cs.flo = curr.flo; cs.fhi = curr.fhi; cs.bnode = curr.bnode; cs.blo = curr.blo; cs.bhi = curr.bhi;
                const BdState walk = go_right ? cs : bd_flip(cs);
                const gb_node_rec crec = load_node(ix, walk.fnode);
                const EdgeFan fan = record_fan(ix, crec, walk.flo, walk.fhi);
                for (uint32_t e = 0; e < fan.n_edges; e++) {
                    uint32_t to; int32_t first, cnt, rev;
                    if (fan.n_edges <= 32) {
                        to = __shfl_sync(FULL, fan.to, e); first = __shfl_sync(FULL, fan.first, e);
                        cnt = __shfl_sync(FULL, fan.cnt, e); rev = __shfl_sync(FULL, fan.rev, e);
                    } else {
                        record_edge_generic(ix, crec, walk.flo, walk.fhi, e, to, first, cnt, rev);
                    }
                    if (to == 0 || cnt <= 0) continue;
                    BdState ns = bd_apply(walk, to, first, cnt, rev);
                    if (!go_right) ns = bd_flip(ns);
                    const uint32_t handle = go_right ? to : (to ^ 1u);      // gbwt_extender.cpp:653
                    const gb_node_rec nrec = load_node(ix, handle);
                    QEntry next = curr;
                    next.fnode = ns.fnode; next.flo = ns.flo; next.fhi = ns.fhi; next.bnode = ns.bnode; next.blo = ns.blo; next.bhi = ns.bhi;
                    uint32_t internal = curr.internal_score;
                    uint32_t used;
                    if (go_right) used = match_fwd(sread, curr.read_hi, ix.seq + nrec.seq_off, 0, min(read_len - curr.read_hi, nrec.len), internal, limit);
                    else used = match_bwd(sread, curr.read_lo, ix.seq + nrec.seq_off, nrec.len, min(curr.read_lo, nrec.len), internal, limit);
                    if (used == 0) continue;               // no base matched within the mismatch budget
                    next.internal_score = internal;
                    if (an >= a_cap || qn + 1 >= q_cap) { status = GB_ITEM_QUEUE_FULL; break; }
                    if (lane == 0) arena[an] = ArenaNode{handle, go_right ? curr.right_tail : curr.left_head};
                    if (go_right) {
                        next.read_hi = curr.read_hi + used;
                        next.right_tail = an++;
                        if (next.read_hi >= read_len) { next.flags |= F_RIGHT_FULL | F_RIGHT_MAX; next.old_score = next.internal_score; }
                        else if (used < nrec.len) { next.flags |= F_RIGHT_MAX; next.old_score = next.internal_score; }
                        num_extensions += (uint32_t)cnt;
                    } else {
                        next.read_lo = curr.read_lo - used; next.offset = nrec.len - used;
                        next.left_head = an++;
                        if (next.read_lo == 0) next.flags |= F_LEFT_FULL | F_LEFT_MAX;
                        else if (next.offset > 0) next.flags |= F_LEFT_MAX;
                        found_extension = true;
                    }
                    set_score(next, p.sc);
                    next.number = number++;
                    q_store(queue + qn, next); qn++;
                }
                if (status != GB_ITEM_OK) break;
                if (go_right) {
                    if (num_extensions < (uint32_t)cs.size()) {
                        curr.flags |= F_RIGHT_MAX; curr.old_score = curr.internal_score; curr.number = number++;
                        if (qn + 1 >= q_cap) { status = GB_ITEM_QUEUE_FULL; break; }
                        q_store(queue + qn, curr); qn++;
                    }
                    __syncwarp();
                    continue;
                }
                __syncwarp();
                if (!found_extension) curr.flags |= F_LEFT_MAX;
                else continue;
            }

            // Case 3: maximal extension (gbwt_extender.cpp:689-691)
            if (!have_best || best.score < curr.score) { best = curr; have_best = true; }
        }
        if (status != GB_ITEM_OK) break;

        // gbwt_extender.cpp:695-700: add the best match to the result
        if (have_best && best.read_hi > best.read_lo) {
            // materialise the path: left chain (already in order) + seed node + right chain reversed
            __syncwarp();
            uint32_t n_left = 0, n_right = 0;
            for (uint32_t a = best.left_head; a != NONE; a = arena[a].parent) n_left++;
            for (uint32_t a = best.right_tail; a != NONE; a = arena[a].parent) n_right++;
            const uint32_t plen = n_left + 1 + n_right;
            if (n_res >= p.max_ext || path_used + plen > p.path_cap) { status = GB_ITEM_OUT_FULL; break; }
            if (lane == 0) {
                uint32_t w = path_used;
                for (uint32_t a = best.left_head; a != NONE; a = arena[a].parent) path_pool[w++] = arena[a].node;
                path_pool[w++] = seed_node;
                uint32_t r = path_used + plen;
                for (uint32_t a = best.right_tail; a != NONE; a = arena[a].parent) path_pool[--r] = arena[a].node;
                gb_extension o;
                o.path_off = path_used; o.path_len = plen; o.mism_off = 0; o.mism_len = 0;
                o.offset = best.offset; o.read_lo = best.read_lo; o.read_hi = best.read_hi; o.score = best.score;
                o.flags = best.flags & (F_LEFT_FULL | F_RIGHT_FULL);
                o.fwd_node = best.fnode; o.fwd_lo = (uint32_t)best.flo; o.fwd_hi = (uint32_t)best.fhi;
                o.bwd_node = best.bnode; o.bwd_lo = (uint32_t)best.blo; o.bwd_hi = (uint32_t)best.bhi;
                o.mismatches = best.internal_score;
                ext[n_res] = o;
            }
            const bool full = (best.flags & (F_LEFT_FULL | F_RIGHT_FULL)) == (F_LEFT_FULL | F_RIGHT_FULL);
            if (full && (best_alignment == NONE || best.internal_score < best_alignment_mm)) {
                best_alignment = n_res; best_alignment_mm = best.internal_score;
            }
            const bool newly_exact = full && best_alignment == n_res && best.internal_score == 0;
            path_used += plen; n_res++;
            __syncwarp();
            if (newly_exact && dead_ok) {
                // GaplessExtension::contains (gbwt_extender.cpp:41-53) for every seed at once: lanes hold the
                // (node, read offset - node offset) of up to 32 path nodes, every lane tests its own seeds
                uint32_t read_offset = best.read_lo;
                for (uint32_t pb = 0; pb < plen; pb += 32) {
                    const uint32_t i = pb + lane;
                    uint32_t h = 0, len = 0;
                    if (i < plen) { h = path_pool[path_used - plen + i]; len = load_node(ix, h).len - (i == 0 ? best.offset : 0u); }
                    const uint32_t incl = (uint32_t)warp_incl_scan((int)len);
                    const int32_t diag = (int32_t)(read_offset + incl - len) - (int32_t)(i == 0 ? best.offset : 0u);
                    const uint32_t cnt = min(32u, plen - pb);
                    for (uint32_t j = lane, c = 0; c * 32 < n_seeds; j += 32, c++) {
                        gb_seed sd; sd.node = 0; sd.diag = 0;
                        if (j < n_seeds) sd = seeds[j];
                        bool hit = false;
                        for (uint32_t x = 0; x < cnt; x++) {
                            const uint32_t hn = __shfl_sync(FULL, h, x); const int32_t dg = __shfl_sync(FULL, diag, x);
                            hit |= sd.node == hn && sd.diag == dg;
                        }
                        if (hit && j < n_seeds) dead |= 1u << c;
                    }
                    read_offset += __shfl_sync(FULL, incl, 31);
                }
            }
        }
    }

    __syncwarp();
    *status_out = status;
    return status == GB_ITEM_OK ? n_res : 0;
}

// A raw record set that extend_finish would leave unchanged: one exact full-length extension.
__device__ __forceinline__ bool extend_result_is_final(const gb_extension* ext, uint32_t n_res) {
    if (n_res == 0) return true;
    if (n_res != 1) return false;
    return (ext[0].flags & 3u) == 3u && ext[0].mismatches == 0;
}

__device__ inline uint32_t extend_finish(const DevIndex& ix, const ExtendParams& p, const uint8_t* sread,
                                         gb_extension* ext, uint32_t n_res, uint32_t* path_pool, uint32_t* mism_pool,
                                         uint32_t* status_out) {
    const int lane = lane_id();
    uint32_t status = GB_ITEM_OK;
    // the best full-length alignment of the search (gbwt_extender.cpp:696-699): fewest mismatches among the full-length records
    uint32_t best_alignment = NONE, best_alignment_mm = 0;
    for (uint32_t i = 0; i < n_res; i++) {
        const uint32_t fl = ext[i].flags, mmc = ext[i].mismatches;
        if ((fl & 3u) == 3u && (best_alignment == NONE || mmc < best_alignment_mm)) { best_alignment = i; best_alignment_mm = mmc; }
    }
    __syncwarp();

    uint32_t mism_used = 0;
    const bool full_length_branch = best_alignment != NONE && best_alignment_mm <= p.max_mismatches;
    if (full_length_branch) {
        // handle_full_length (gbwt_extender.cpp:301-329), stable
        if (lane == 0) {
            for (uint32_t i = 1; i < n_res; i++) {
                gb_extension key = ext[i];
                const bool kfull = (key.flags & 3u) == 3u;
                uint32_t j = i;
                while (j > 0) {
                    const gb_extension& o = ext[j - 1];
                    const bool ofull = (o.flags & 3u) == 3u;
                    const bool less = (kfull && ofull) ? (key.mismatches < o.mismatches) : (kfull && !ofull);
                    if (!less) break;
                    ext[j] = ext[j - 1]; j--;
                }
                ext[j] = key;
            }
            uint32_t tail = 0;
            for (uint32_t i = 0; i < n_res; i++) {
                if ((ext[i].flags & 3u) != 3u) break;
                bool overlap = false;
                for (uint32_t prev = 0; prev < tail; prev++) {
                    const uint32_t ov = ext_overlap(ix, ext[i], ext[prev], path_pool);
                    if ((double)ov > p.overlap_threshold * (double)(ext[prev].read_hi - ext[prev].read_lo)) { overlap = true; break; }
                }
                if (overlap) continue;
                if (i > tail) ext[tail] = ext[i];
                tail++;
            }
            n_res = tail;
        }
        n_res = __shfl_sync(FULL, n_res, 0);
        __syncwarp();
    } else {
        n_res = remove_duplicates(ext, n_res);
    }
    // find_mismatches for every surviving extension (one call site for both branches)
    for (uint32_t i = 0; i < n_res; i++) {
        gb_extension e = ext[i];
        if (!find_mismatches(ix, e, sread, path_pool, mism_pool, mism_used, p.mism_cap)) { status = GB_ITEM_OUT_FULL; break; }
        __syncwarp();
        if (lane == 0) ext[i] = e;
    }
    if (!full_length_branch) {
        __syncwarp();
        if (status == GB_ITEM_OK && p.trim) {
            bool trimmed = false;
            for (uint32_t i = 0; i < n_res; i++) {
                gb_extension e = ext[i];
                const bool t = trim_mismatches(ix, e, p.sc, path_pool, mism_pool);
                if (t) { trimmed = true; if (lane == 0) ext[i] = e; }
                __syncwarp();
            }
            if (trimmed) n_res = remove_duplicates(ext, n_res);
        }
    }
    __syncwarp();
    if (status == GB_ITEM_OK && lane == 0) {
        // `mismatches` reports the final mismatch count (internal_score is stale after trimming)
        for (uint32_t i = 0; i < n_res; i++) ext[i].mismatches = ext[i].mism_len;
    }
    __syncwarp();
    *status_out = status;
    return status == GB_ITEM_OK ? n_res : 0;
}

__device__ inline uint32_t extend_item(const DevIndex& ix, const ExtendParams& p,
                                       const uint8_t* sread, uint32_t read_len,
                                       const gb_seed* seeds, uint32_t n_seeds,
                                       QEntry* queue, uint32_t q_cap, ArenaNode* arena, uint32_t a_cap,
                                       gb_extension* ext, uint32_t* path_pool, uint32_t* mism_pool,
                                       uint32_t* status_out) {
    const uint32_t n = extend_search(ix, p, sread, read_len, seeds, n_seeds, queue, q_cap, arena, a_cap, ext, path_pool, status_out);
    if (*status_out != GB_ITEM_OK || extend_result_is_final(ext, n)) return n;
    __syncwarp();
    return extend_finish(ix, p, sread, ext, n, path_pool, mism_pool, status_out);
}

} // namespace gb
