// seed.cuh — seed_kernel: one warp per read does
//   find_minimizers            minimizer_mapper.cpp:3918-3974  (gbwtgraph minimizer_regions + find)
//   sort_minimizers_by_score   :4074-4107  (LazyRNG tie shuffle, utility.hpp:771-794)
//   find_seeds                 :4109-4517  (filter cascade with running state)
//   cluster_seeds              snarl_seed_clusterer.cpp:28-63 (connected components under the
//                              unoriented minimum distance, via the 16-byte distance payload)
//   score_cluster              :4738-4780
//   cluster selection          :680-832    (process_until_threshold_e + cluster score cutoff)
//   extend_seed_group packing  :4784-4850  ((handle, read_offset - node_offset) seeds)
// and leaves DevMinimizer / DevSeed / DevItem records in HBM for the next kernels.
#pragma once
#include "map_state.cuh"
#include "minimizer_common.h"

namespace gb {

constexpr uint32_t MAX_CLUSTERS = 64;     // clusters per read kept in shared memory

struct SeedPools {
    DevMinimizer* minimizers; uint32_t min_cap;  uint32_t* min_cursor;
    DevSeed* seeds;           uint32_t seed_cap; uint32_t* seed_cursor;
    DevItem* items;           uint32_t item_cap; uint32_t* item_cursor;
    gb_seed* ext_seeds;       uint32_t ext_cap;  uint32_t* ext_cursor;
};

// Per-warp shared memory carve-up.
struct SeedSmem {
    uint8_t*  read;        // [Lc]
    uint64_t* khash;       // [Lc]  canonical hash per k-mer start (scratch after compaction)
    uint64_t* kkey;        // [Lc]  canonical key
    uint8_t*  kflag;       // [Lc]  bit0 valid, bit1 reverse
    uint64_t* m_key;       // [MAX_MINIMIZERS]
    uint64_t* m_hash;
    double*   m_score;
    uint32_t* m_hit_off;
    uint32_t* m_hit_cnt;
    uint16_t* m_fwd;
    uint16_t* m_agg_start;
    uint16_t* m_agg_len;
    uint8_t*  m_rev;
    uint8_t*  m_order;     // score order -> read order index
    uint8_t*  m_pass;      // per score-order minimizer: passed the filters
    // clusters
    double*   c_score;     // [MAX_CLUSTERS]
    double*   c_cov;
    uint32_t* c_label;
    uint32_t* c_present;   // [MAX_CLUSTERS * PRESENT_WORDS]
    uint8_t*  c_order;     // processing order
};

__host__ __device__ inline size_t seed_smem_bytes(uint32_t Lc) {
    size_t b = 0;
    b += (size_t)Lc * 8 * 2;                       // khash, kkey
    b += (size_t)MAX_MINIMIZERS * (8 + 8 + 8);     // m_key, m_hash, m_score
    b += (size_t)MAX_CLUSTERS * (8 + 8);           // c_score, c_cov
    b += (size_t)MAX_MINIMIZERS * (4 + 4);         // hit_off, hit_cnt
    b += (size_t)MAX_CLUSTERS * 4 * (1 + PRESENT_WORDS);
    b += (size_t)MAX_MINIMIZERS * (2 + 2 + 2);     // fwd, agg_start, agg_len
    b += (size_t)Lc * 2;                           // read, kflag
    b += (size_t)MAX_MINIMIZERS * 3;               // rev, order, pass
    b += MAX_CLUSTERS;                             // c_order
    return (b + 15) & ~(size_t)15;
}

__device__ inline SeedSmem carve_seed_smem(uint8_t* base, uint32_t Lc) {
    SeedSmem s;
    uint8_t* p = base;
    s.khash = (uint64_t*)p; p += (size_t)Lc * 8;
    s.kkey = (uint64_t*)p; p += (size_t)Lc * 8;
    s.m_key = (uint64_t*)p; p += MAX_MINIMIZERS * 8;
    s.m_hash = (uint64_t*)p; p += MAX_MINIMIZERS * 8;
    s.m_score = (double*)p; p += MAX_MINIMIZERS * 8;
    s.c_score = (double*)p; p += MAX_CLUSTERS * 8;
    s.c_cov = (double*)p; p += MAX_CLUSTERS * 8;
    s.m_hit_off = (uint32_t*)p; p += MAX_MINIMIZERS * 4;
    s.m_hit_cnt = (uint32_t*)p; p += MAX_MINIMIZERS * 4;
    s.c_label = (uint32_t*)p; p += MAX_CLUSTERS * 4;
    s.c_present = (uint32_t*)p; p += MAX_CLUSTERS * 4 * PRESENT_WORDS;
    s.m_fwd = (uint16_t*)p; p += MAX_MINIMIZERS * 2;
    s.m_agg_start = (uint16_t*)p; p += MAX_MINIMIZERS * 2;
    s.m_agg_len = (uint16_t*)p; p += MAX_MINIMIZERS * 2;
    s.read = p; p += Lc;
    s.kflag = p; p += Lc;
    s.m_rev = p; p += MAX_MINIMIZERS;
    s.m_order = p; p += MAX_MINIMIZERS;
    s.m_pass = p; p += MAX_MINIMIZERS;
    s.c_order = p; p += MAX_CLUSTERS;
    return s;
}

__device__ __forceinline__ uint32_t pow13(uint32_t e) {
    uint32_t r = 1, b = 13;
    while (e) { if (e & 1) r *= b; b *= b; e >>= 1; }
    return r;
}

// minimum graph distance a -> b on the forward strand (see gb_dist_payload)
__device__ __forceinline__ bool seeds_within(const DevSeed& a, const DevSeed& b, int32_t limit) {
    const uint32_t ida = a.id_off >> 10, idb = b.id_off >> 10;
    if (ida == idb) {
        const int32_t d = (int32_t)(b.id_off & 1023u) - (int32_t)(a.id_off & 1023u);
        return (d >= 0 ? d : -d) <= limit;
    }
    if (a.slot < b.slot) return (b.c_in - a.c_out) <= limit;
    if (b.slot < a.slot) return (a.c_in - b.c_out) <= limit;
    return false;
}

// The whole seeding stage for one read.  Returns status.
__device__ inline uint32_t seed_read(const DevIndex& ix, const MapParamsDev& P, const SeedSmem& sm,
                                     const uint8_t* __restrict__ gread, uint32_t L, uint32_t read_idx,
                                     const SeedPools& pools, ReadState& rs) {
    const int lane = lane_id();
    const uint32_t k = ix.k, w = ix.w;
    const uint32_t window_bp = k + w - 1;
    rs.min_off = rs.min_cnt = rs.item_off = rs.item_cnt = rs.seed_off = rs.seed_cnt = 0; rs.n_clusters = 0;
    rs.rng.inited = 0; rs.rng.state = 0;

    // ---- stage the read; fold the RNG seed: seed = seed * 13 + byte over the sequence -------
    {
        const uint32_t chunk = (L + 31) / 32;
        const uint32_t b = min(L, lane * chunk), e = min(L, b + chunk);
        uint32_t fold = 0;
        for (uint32_t i = b; i < e; i++) { const uint8_t c = gread[i]; sm.read[i] = c; fold = fold * 13u + c; }
        uint32_t term = fold * pow13(L - e);       // bytes after my chunk
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) term += __shfl_xor_sync(FULL, term, o);
        rs.rng.seed = term;
    }
    __syncwarp();
    if (L > 512) return GB_ITEM_OUT_FULL;          // short-read path: coverage bitmaps are 512 bits
    if (L < window_bp) return GB_ITEM_OK;          // no minimizers -> no seeds -> unmapped

    // ---- canonical k-mer hashes ------------------------------------------------------------
    const uint32_t nk = L - k + 1;
    {
        const uint64_t mask = (1ull << (2 * k)) - 1ull;
        const uint32_t chunk = (nk + 31) / 32;
        const uint32_t s0 = min(nk, lane * chunk), s1 = min(nk, s0 + chunk);
        if (s0 < s1) {
            uint64_t fk = 0, rk = 0; uint32_t run = 0;
            for (uint32_t i = s0; i < s1 + k - 1; i++) {
                const uint32_t c = gbmin::base_code(sm.read[i]);
                if (c < 4) { fk = ((fk << 2) | c) & mask; rk = (rk >> 2) | ((uint64_t)(3 - c) << (2 * (k - 1))); run++; }
                else { fk = 0; rk = 0; run = 0; }
                if (i + 1 >= s0 + k) {
                    const uint32_t s = i + 1 - k;
                    uint8_t flag = 0; uint64_t h = ~0ull, key = 0;
                    if (run >= k) {
                        const uint64_t hf = gbmin::hash64(fk), hr = gbmin::hash64(rk);
                        if (hr < hf) { h = hr; key = rk; flag = 3; } else { h = hf; key = fk; flag = 1; }
                    }
                    sm.khash[s] = h; sm.kkey[s] = key; sm.kflag[s] = flag;
                }
            }
        }
    }
    __syncwarp();

    // ---- window minima -> minimizers in read order ---------------------------------------------
    uint32_t M = 0;
    {
        const int32_t last_window = (int32_t)(L - window_bp);
        for (uint32_t base = 0; base < nk; base += 32) {
            const uint32_t s = base + lane;
            bool is_min = false; int32_t lo = 0, hi = -1;
            if (s < nk && (sm.kflag[s] & 1)) {
                const uint64_t h = sm.khash[s];
                int32_t l = -1000000, r = 1000000;
                for (int32_t t = (int32_t)s - 1; t >= 0 && t > (int32_t)s - (int32_t)w; t--)
                    if ((sm.kflag[t] & 1) && sm.khash[t] < h) { l = t; break; }
                for (int32_t t = (int32_t)s + 1; t < (int32_t)nk && t < (int32_t)s + (int32_t)w; t++)
                    if ((sm.kflag[t] & 1) && sm.khash[t] < h) { r = t; break; }
                lo = max(max((int32_t)s - (int32_t)w + 1, l + 1), 0);
                hi = min(min((int32_t)s, r - (int32_t)w), last_window);
                is_min = lo <= hi;
            }
            const uint32_t bal = __ballot_sync(FULL, is_min);
            const uint32_t cnt = __popc(bal);
            if (M + cnt > MAX_MINIMIZERS) return GB_ITEM_OUT_FULL;
            if (is_min) {
                const uint32_t idx = M + __popc(bal & ((1u << lane) - 1u));
                sm.m_key[idx] = sm.kkey[s]; sm.m_hash[idx] = sm.khash[s];
                sm.m_fwd[idx] = (uint16_t)s; sm.m_rev[idx] = (sm.kflag[s] >> 1) & 1;
                sm.m_agg_start[idx] = (uint16_t)lo; sm.m_agg_len[idx] = (uint16_t)(hi - lo + (int32_t)window_bp);
            }
            M += cnt;
        }
    }
    __syncwarp();
    if (M == 0) return GB_ITEM_OK;

    // ---- index lookup + score ----------------------------------------------------------------------
    for (uint32_t a = lane; a < M; a += 32) {
        const uint64_t key = sm.m_key[a];
        uint64_t h = gbmin::hash64(key) & ix.table_mask;
        uint32_t off = 0, cnt = 0;
        while (true) {
            const uint4 cell = __ldg(reinterpret_cast<const uint4*>(ix.table) + h);
            const uint64_t ckey = ((uint64_t)cell.y << 32) | cell.x;
            if (ckey == GB_NO_KEY) break;
            if (ckey == key) { off = cell.z; cnt = cell.w; break; }
            h = (h + 1) & ix.table_mask;
        }
        sm.m_hit_off[a] = off; sm.m_hit_cnt[a] = cnt;
        sm.m_score[a] = cnt == 0 ? 0.0 : (cnt <= P.hard_hit_cap ? P.hit_score_table[cnt] : 1.0);
    }
    __syncwarp();

    // ---- score order (stable rank sort on (score desc, key asc)) -----------------------------------
    for (uint32_t a = lane; a < M; a += 32) {
        const double sa = sm.m_score[a]; const uint64_t ka = sm.m_key[a];
        uint32_t rank = 0;
        for (uint32_t b = 0; b < M; b++) {
            const double sb = sm.m_score[b]; const uint64_t kb = sm.m_key[b];
            const bool before = sb > sa || (sb == sa && (kb < ka || (kb == ka && b < a)));
            rank += before ? 1u : 0u;
        }
        sm.m_order[rank] = (uint8_t)a;
    }
    __syncwarp();

    // ---- shuffle the runs tied at the top score (sort_shuffling_ties over runs) ----------------------
    // khash/kkey are dead now: reuse as scratch.
    uint8_t* run_begin = reinterpret_cast<uint8_t*>(sm.khash);          // [<= M]
    uint8_t* run_len = run_begin + MAX_MINIMIZERS;
    uint8_t* tmp_order = run_len + MAX_MINIMIZERS;
    DevRng rng = rs.rng;
    if (lane == 0) {
        const double top = sm.m_score[sm.m_order[0]];
        uint32_t T = 0, pos = 0;
        while (pos < M && sm.m_score[sm.m_order[pos]] == top) {
            uint32_t e = pos + 1;
            while (e < M && sm.m_key[sm.m_order[e]] == sm.m_key[sm.m_order[pos]]) e++;
            run_begin[T] = (uint8_t)pos; run_len[T] = (uint8_t)(e - pos); T++;
            pos = e;
        }
        const uint32_t tied_end = pos;
        if (T > 1) {
            for (uint32_t i = 1; i < T; i++) {
                const uint32_t j = rng_next(rng) % (i + 1);
                const uint8_t tb = run_begin[j], tl = run_len[j];
                run_begin[j] = run_begin[i]; run_len[j] = run_len[i];
                run_begin[i] = tb; run_len[i] = tl;
            }
            uint32_t wpos = 0;
            for (uint32_t t = 0; t < T; t++)
                for (uint32_t x = 0; x < run_len[t]; x++) tmp_order[wpos++] = sm.m_order[run_begin[t] + x];
            for (uint32_t x = 0; x < tied_end; x++) sm.m_order[x] = tmp_order[x];
        }
    }
    rng.state = __shfl_sync(FULL, rng.state, 0); rng.inited = __shfl_sync(FULL, rng.inited, 0);
    __syncwarp();

    // ---- find_seeds filter cascade (sequential running state; lane 0) -------------------------------------
    uint32_t total_hits = 0;
    if (lane == 0) {
        double base_target_score = 0.0, target_score = 0.0, selected_score = 0.0;
        const bool use_fraction = (P.hit_cap != 0 || P.minimizer_score_fraction != 1.0);
        if (use_fraction) {
            for (uint32_t i = 0; i < M; i++) base_target_score += sm.m_score[sm.m_order[i]];
            target_score = (base_target_score * P.minimizer_score_fraction) + 0.000001;
        }
        uint32_t limit = 0, run_hits = 0; bool taking_run = false;
        uint32_t num_minimizers = 0, worst_kept_hits = 0;
        const uint32_t num_min_by_read_len = L / P.num_bp_per_min;
        // read_coverage bit vector (only consulted once num_minimizers reaches the cap)
        uint32_t* cov = reinterpret_cast<uint32_t*>(sm.kkey);
        const uint32_t cov_words = (L + 31) / 32;
        for (uint32_t x = 0; x < cov_words; x++) cov[x] = 0;
        for (uint32_t i = 0; i < M; i++) {
            const uint32_t a = sm.m_order[i];
            if (i >= limit) {
                limit = i + 1; run_hits = sm.m_hit_cnt[a];
                for (uint32_t j = i + 1; j < M && sm.m_key[sm.m_order[j]] == sm.m_key[a]; j++) { limit++; run_hits += sm.m_hit_cnt[sm.m_order[j]]; }
                taking_run = false;
            }
            const uint32_t hits = sm.m_hit_cnt[a];
            const double score = sm.m_score[a];
            bool passing = hits > 0;                                        // any-hits
            if (passing) passing = run_hits <= P.hard_hit_cap;             // hard-hit-cap
            if (passing && P.max_unique_min != 0) {                        // max-min||num-bp-per-min
                const uint32_t fwd = sm.m_fwd[a];
                const uint32_t cs = fwd < P.minimizer_coverage_flank ? 0 : fwd - P.minimizer_coverage_flank;
                const uint32_t ce = min(L, fwd + k + P.minimizer_coverage_flank);
                if (num_minimizers < max(P.max_unique_min, num_min_by_read_len)) {
                    for (uint32_t x = cs; x < ce; x++) cov[x >> 5] |= 1u << (x & 31);
                    worst_kept_hits = max(hits, worst_kept_hits);
                } else if (hits > worst_kept_hits) {
                    passing = false;
                } else {
                    bool covered = false;
                    for (uint32_t x = cs; x < ce; x++) if (cov[x >> 5] & (1u << (x & 31))) { covered = true; break; }
                    if (covered) passing = false;
                    else for (uint32_t x = cs; x < ce; x++) cov[x >> 5] |= 1u << (x & 31);
                }
            }
            if (passing && use_fraction) {                                  // hit-cap||score-fraction
                passing = (hits <= P.hit_cap) || (run_hits <= P.hard_hit_cap && selected_score + score <= target_score) || taking_run;
                if (passing) selected_score += score; else target_score = selected_score;
            }
            sm.m_pass[i] = passing ? 1 : 0;
            if (passing) { taking_run = true; num_minimizers++; total_hits += hits; }
        }
    }
    total_hits = __shfl_sync(FULL, total_hits, 0);
    __syncwarp();

    // ---- minimizer records (score order) for the MAPQ cap ------------------------------------------------
    uint32_t min_off = 0;
    if (lane == 0) min_off = atomicAdd(pools.min_cursor, M);
    min_off = __shfl_sync(FULL, min_off, 0);
    if (min_off + M > pools.min_cap) return GB_ITEM_OUT_FULL;
    for (uint32_t i = lane; i < M; i += 32) {
        const uint32_t a = sm.m_order[i];
        DevMinimizer dm; dm.hash = sm.m_hash[a]; dm.fwd_offset = sm.m_fwd[a]; dm.agg_start = sm.m_agg_start[a]; dm.agg_len = sm.m_agg_len[a]; dm.pad = 0;
        pools.minimizers[min_off + i] = dm;
    }
    rs.min_off = min_off; rs.min_cnt = M;
    rs.rng = rng;
    if (total_hits == 0) return GB_ITEM_OK;

    // ---- seeds ----------------------------------------------------------------------------------------------
    uint32_t seed_off = 0;
    if (lane == 0) seed_off = atomicAdd(pools.seed_cursor, total_hits);
    seed_off = __shfl_sync(FULL, seed_off, 0);
    if (seed_off + total_hits > pools.seed_cap) return GB_ITEM_OUT_FULL;
    DevSeed* seeds = pools.seeds + seed_off;
    {
        uint32_t wpos = 0;
        for (uint32_t i = 0; i < M; i++) {
            if (!sm.m_pass[i]) continue;
            const uint32_t a = sm.m_order[i];
            const uint32_t hits = sm.m_hit_cnt[a], hoff = sm.m_hit_off[a];
            const bool rev = sm.m_rev[a];
            for (uint32_t j = lane; j < hits; j += 32) {
                const gb_hit* hp = ix.hits + hoff + j;
                // gb_hit is 24 bytes (8-byte aligned): three 64-bit loads
                const uint2 pw = __ldg(reinterpret_cast<const uint2*>(hp));
                const uint2 p1 = __ldg(reinterpret_cast<const uint2*>(hp) + 1);
                const uint2 p2 = __ldg(reinterpret_cast<const uint2*>(hp) + 2);
                uint4 pl; pl.x = p1.x; pl.y = p1.y; pl.z = p2.x; pl.w = p2.y;
                const uint64_t pos = ((uint64_t)pw.y << 32) | pw.x;
                uint32_t node = (uint32_t)(pos >> 10), off = (uint32_t)(pos & 1023u);
                const uint32_t nlen = load_node(ix, node).len;
                if (rev) { node ^= 1u; off = nlen - off - 1; }          // reverse_base_pos, :4462-4465
                const uint32_t off_f = (node & 1u) ? nlen - 1 - off : off;
                DevSeed s;
                s.node = node; s.offset = off; s.source = i; s.label = wpos + j;
                s.c_in = (int32_t)pl.x + (int32_t)off_f;
                s.c_out = (int32_t)pl.y - (int32_t)(nlen - off_f);
                s.slot = pl.z;
                s.id_off = ((node >> 1) << 10) | off_f;
                seeds[wpos + j] = s;
            }
            wpos += hits;
        }
    }
    rs.seed_off = seed_off; rs.seed_cnt = total_hits;
    __syncwarp();

    // ---- clustering: label propagation to the smallest seed index of each component ---------------------------
    const uint32_t H = total_hits;
    const int32_t limit = (int32_t)max(P.distance_limit, L + 50);      // get_distance_limit, minimizer_mapper.hpp:554
    while (true) {
        bool changed = false;
        for (uint32_t i = lane; i < H; i += 32) {
            const DevSeed si = seeds[i];
            uint32_t best = si.label;
            for (uint32_t j = 0; j < H; j++) {
                if (j == i) continue;
                const DevSeed sj = seeds[j];
                if (sj.label < best && seeds_within(si, sj, limit)) best = sj.label;
            }
            if (best != si.label) { seeds[i].label = best; changed = true; }
        }
        __syncwarp();
        if (!__any_sync(FULL, changed)) break;
    }

    // ---- clusters in order of their first seed; score_cluster ---------------------------------------------------
    uint32_t C = 0;
    for (uint32_t base = 0; base < H; base += 32) {
        const uint32_t i = base + lane;
        const bool root = i < H && seeds[i].label == i;
        const uint32_t bal = __ballot_sync(FULL, root);
        if (C + __popc(bal) > MAX_CLUSTERS) return GB_ITEM_OUT_FULL;
        if (root) sm.c_label[C + __popc(bal & ((1u << lane) - 1u))] = i;
        C += __popc(bal);
    }
    __syncwarp();
    rs.n_clusters = C;
    for (uint32_t c = lane; c < C; c += 32) {
        const uint32_t label = sm.c_label[c];
        uint32_t present[PRESENT_WORDS];
#pragma unroll
        for (uint32_t x = 0; x < PRESENT_WORDS; x++) present[x] = 0;
        for (uint32_t i = 0; i < H; i++) if (seeds[i].label == label) { const uint32_t s = seeds[i].source; present[s >> 5] |= 1u << (s & 31); }
        double score = 0.0;
        uint32_t covered[16];                     // up to 512 bp
#pragma unroll
        for (uint32_t x = 0; x < 16; x++) covered[x] = 0;
        for (uint32_t j = 0; j < M; j++) {
            if (!(present[j >> 5] & (1u << (j & 31)))) continue;
            const uint32_t a = sm.m_order[j];
            score += sm.m_score[a];
            const uint32_t s0 = sm.m_fwd[a];
            for (uint32_t x = s0; x < s0 + k && x < L; x++) covered[x >> 5] |= 1u << (x & 31);
        }
        uint32_t cnt = 0;
#pragma unroll
        for (uint32_t x = 0; x < 16; x++) cnt += __popc(covered[x]);
        sm.c_score[c] = score;
        sm.c_cov[c] = (double)cnt / (double)L;
#pragma unroll
        for (uint32_t x = 0; x < PRESENT_WORDS; x++) sm.c_present[c * PRESENT_WORDS + x] = present[x];
    }
    __syncwarp();

    // ---- cluster selection (minimizer_mapper.cpp:655-832) --------------------------------------------------------------
    uint32_t n_kept = 0;
    uint8_t* kept = tmp_order;                    // scratch: kept cluster ids in processing order
    if (lane == 0) {
        double best_cluster_score = 0.0, second_best_cluster_score = 0.0;
        for (uint32_t c = 0; c < C; c++) {
            const double sc = sm.c_score[c];
            if (sc > best_cluster_score) { second_best_cluster_score = best_cluster_score; best_cluster_score = sc; }
            else if (sc > second_best_cluster_score) second_best_cluster_score = sc;
        }
        double cluster_score_cutoff = best_cluster_score - P.cluster_score_threshold;
        if (cluster_score_cutoff - P.pad_cluster_score_threshold < second_best_cluster_score)
            cluster_score_cutoff = min(cluster_score_cutoff, second_best_cluster_score);
        // sort_shuffling_ties with comparator (coverage desc, then score desc); stable insertion sort
        auto comes_before = [&](uint32_t a, uint32_t b) {
            return (sm.c_cov[a] > sm.c_cov[b]) || (sm.c_cov[a] == sm.c_cov[b] && sm.c_score[a] > sm.c_score[b]);
        };
        for (uint32_t c = 0; c < C; c++) {
            uint32_t j = c;
            while (j > 0 && comes_before(c, sm.c_order[j - 1])) { sm.c_order[j] = sm.c_order[j - 1]; j--; }
            sm.c_order[j] = (uint8_t)c;
        }
        uint32_t ties = 0;
        while (ties < C && !comes_before(sm.c_order[0], sm.c_order[ties])) ties++;
        for (uint32_t i = 1; i < ties; i++) {
            const uint32_t j = rng_next(rng) % (i + 1);
            const uint8_t t = sm.c_order[j]; sm.c_order[j] = sm.c_order[i]; sm.c_order[i] = t;
        }
        // process_until_threshold_e, minimizer_mapper.hpp:1617-1657
        const double cutoff = C == 0 ? 0.0 : sm.c_cov[sm.c_order[0]] - P.cluster_coverage_threshold;
        uint32_t unskipped = 0, kept_cluster_count = 0;
        for (uint32_t i = 0; i < C; i++) {
            const uint32_t c = sm.c_order[i];
            bool process = false;
            if (P.cluster_coverage_threshold != 0 && sm.c_cov[c] <= cutoff) process = unskipped < P.min_extensions;
            else process = unskipped < P.max_extensions;
            if (!process) continue;
            // additional score filter (:746-762); escaped_threshold is always false here
            if (P.cluster_score_threshold != 0 && sm.c_score[c] < cluster_score_cutoff && kept_cluster_count >= P.min_extensions) continue;
            kept[n_kept++] = (uint8_t)c;
            kept_cluster_count++; unskipped++;
        }
    }
    n_kept = __shfl_sync(FULL, n_kept, 0);
    rng.state = __shfl_sync(FULL, rng.state, 0); rng.inited = __shfl_sync(FULL, rng.inited, 0);
    rs.rng = rng;
    __syncwarp();
    if (n_kept == 0) return GB_ITEM_OK;

    // ---- work items: (handle, read_offset - node_offset) seeds per kept cluster ------------------------------------------
    uint32_t item_off = 0;
    if (lane == 0) item_off = atomicAdd(pools.item_cursor, n_kept);
    item_off = __shfl_sync(FULL, item_off, 0);
    if (item_off + n_kept > pools.item_cap) return GB_ITEM_OUT_FULL;
    for (uint32_t t = 0; t < n_kept; t++) {
        const uint32_t c = kept[t];
        const uint32_t label = sm.c_label[c];
        // count + reserve
        uint32_t cnt = 0;
        for (uint32_t i = lane; i < H; i += 32) cnt += seeds[i].label == label ? 1u : 0u;
        cnt = (uint32_t)warp_sum((int)cnt);
        uint32_t eoff = 0;
        if (lane == 0) eoff = atomicAdd(pools.ext_cursor, cnt);
        eoff = __shfl_sync(FULL, eoff, 0);
        if (eoff + cnt > pools.ext_cap) return GB_ITEM_OUT_FULL;
        uint32_t wpos = 0;
        for (uint32_t base = 0; base < H; base += 32) {
            const uint32_t i = base + lane;
            const bool mine = i < H && seeds[i].label == label;
            const uint32_t bal = __ballot_sync(FULL, mine);
            if (mine) {
                const DevSeed s = seeds[i];
                const uint32_t a = sm.m_order[s.source];
                const int32_t pin = (int32_t)sm.m_fwd[a] + (sm.m_rev[a] ? (int32_t)k - 1 : 0);   // value.offset
                gb_seed g; g.node = s.node; g.diag = pin - (int32_t)s.offset;                  // to_seed, gbwt_extender.hpp:159
                pools.ext_seeds[eoff + wpos + __popc(bal & ((1u << lane) - 1u))] = g;
            }
            wpos += __popc(bal);
        }
        if (lane == 0) {
            DevItem it; it.read = read_idx; it.seed_off = eoff; it.seed_cnt = cnt; it.pad = 0;
#pragma unroll
            for (uint32_t x = 0; x < PRESENT_WORDS; x++) it.present[x] = sm.c_present[c * PRESENT_WORDS + x];
            pools.items[item_off + t] = it;
        }
    }
    rs.item_off = item_off; rs.item_cnt = n_kept;
    return GB_ITEM_OK;
}

} // namespace gb
