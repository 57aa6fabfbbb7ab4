// tail_plan.cuh — tail_plan_kernel: one warp per slow unit (read or pair) walks the unit's extension sets the way
// align_sets will (minimizer_mapper.cpp:886-1063 / :1904-2043, find_optimal_tail_alignments :5369-5622), builds the
// haplotype forest of every tail that can be asked for (get_tail_forest / dfs_gbwt :5745-6013) and writes one DP tile per
// tree (xdrop_tile.cuh) plus a plan entry per tail.  Included by map.cu inside namespace gb, after align_read.cuh.
//
// The walk is a superset of what align_sets will do: it does not know the LazyRNG state (tie shuffles only reorder work)
// nor the running winner (the score-estimate skip of :5478-5492 can only drop tails), so every tail align_sets can reach
// is planned, and a few it will not reach are computed for nothing.  Anything that does not fit (plan entries, tile pool,
// tile budgets, int16 ranges) is simply left out and aligned in place by align_tail.
#pragma once

struct PlanPools {
    TailPlanEntry* entries; uint32_t* unit_base; uint32_t* unit_count;
    uint8_t* tiles; uint32_t tile_units_cap; uint32_t* tile_units_cursor;      // 16-byte units
    uint32_t* tile_off; uint32_t tile_cap; uint32_t* tile_cursor;
    TileResult* results;
    uint32_t* lists[2 * TILE_CLASSES]; uint32_t* list_count;                   // [2 waves][TILE_CLASSES]
    uint64_t* stats;                                                           // [4] tails planned, trees, tiles, cells upper bound
};

// One tail of one extension: forest -> tiles + entry.  Returns false when the tail is left to align_tail.
__device__ inline bool plan_tail(const DevIndex& ix, const MapParamsDev& P, const DevScores& sc, const TailWs& ws, const gb_extension& e,
                                 const uint32_t* path_pool, const uint8_t* read, uint32_t L, bool left_tail, uint32_t key, uint32_t wave,
                                 const PlanPools& pp, TailPlanEntry* unit_entries, uint32_t& n_entries) {
    const int lane = lane_id();
    if (n_entries >= PLAN_PER_UNIT) return false;
    uint32_t from_node, from_offset, tail_length; int32_t lo, hi;
    const uint32_t first = path_pool[e.path_off], last = path_pool[e.path_off + e.path_len - 1];
    if (left_tail) {
        from_node = first ^ 1u; from_offset = load_node(ix, first).len - e.offset;
        lo = (int32_t)e.bwd_lo; hi = (int32_t)e.bwd_hi; tail_length = e.read_lo;
    } else {
        from_node = last;
        uint32_t tail_off = e.offset + (e.read_hi - e.read_lo);
#pragma unroll 1
        for (uint32_t i = 0; i + 1 < e.path_len; i++) tail_off -= load_node(ix, path_pool[e.path_off + i]).len;
        from_offset = tail_off; lo = (int32_t)e.fwd_lo; hi = (int32_t)e.fwd_hi; tail_length = L - e.read_hi;
    }
    if (tail_length == 0) return false;
    const uint32_t gap = longest_detectable_gap(sc, L, tail_length);
    const uint32_t g = max(gap, 1u);
    if (!tile_scores_fit_int16(sc, tail_length, g)) return false;
    uint32_t root_trim = 0;
    const uint32_t n_forest = build_tail_forest(ix, ws, from_node, lo, hi, from_offset, gap + tail_length, root_trim);
    if (n_forest == 0xffffffffu) return false;
    // trees of the forest: sizes, eligibility, space
    uint32_t n_trees = 0, units = 0;
    {
        uint32_t t0 = 0;
#pragma unroll 1
        while (t0 < n_forest) {
            uint32_t t1 = t0 + 1;
#pragma unroll 1
            while (t1 < n_forest && ws.tree[t1].parent >= 0) t1++;
            uint32_t bases = 0, depth = 0;
#pragma unroll 1
            for (uint32_t i = t0 + lane; i < t1; i += 32) { bases += ws.tree[i].len; depth = max(depth, ws.tree[i].depth); }
            bases = (uint32_t)warp_sum((int)bases); depth = __reduce_max_sync(FULL, depth);
            if ((uint64_t)bases * tail_length <= P.max_dozeu_cells) {
                if (!tile_eligible(sc, tail_length, g, t1 - t0, bases, depth)) return false;
                units += tile_bytes(t1 - t0, bases, tail_length) / 16;
            }
            n_trees++;
            t0 = t1;
        }
    }
    uint32_t first_tile = 0, first_unit = 0;
    if (lane == 0) {
        first_tile = atomicAdd(pp.tile_cursor, n_trees);
        first_unit = atomicAdd(pp.tile_units_cursor, units);
        if (first_tile > pp.tile_cap || n_trees > pp.tile_cap - first_tile || first_unit > pp.tile_units_cap || units > pp.tile_units_cap - first_unit) first_tile = 0xffffffffu;
    }
    first_tile = __shfl_sync(FULL, first_tile, 0); first_unit = __shfl_sync(FULL, first_unit, 0);
    if (first_tile == 0xffffffffu) return false;
    uint32_t t0 = 0, ti = first_tile, unit_at = first_unit;
#pragma unroll 1
    while (t0 < n_forest) {
        uint32_t t1 = t0 + 1;
#pragma unroll 1
        while (t1 < n_forest && ws.tree[t1].parent >= 0) t1++;
        uint32_t bases = 0;
#pragma unroll 1
        for (uint32_t i = t0 + lane; i < t1; i += 32) bases += ws.tree[i].len;
        bases = (uint32_t)warp_sum((int)bases);
        if ((uint64_t)bases * tail_length > P.max_dozeu_cells) {
            if (lane == 0) pp.tile_off[ti] = TILE_REFUSED;
        } else {
            const uint32_t nt = t1 - t0;
            uint8_t* tile = pp.tiles + (size_t)unit_at * 16;
            TileNode* tn = reinterpret_cast<TileNode*>(tile + 32);
            uint8_t* tb = tile + 32 + 8 * ((nt + 1u) & ~1u);
            uint8_t* tq = tb + ((bases + 15u) & ~15u);
            // node table + bases: lane-parallel over nodes for the table, node by node for the bases (nodes are <= 32 bp on these graphs)
            uint32_t at = 0;
#pragma unroll 1
            for (uint32_t i = t0; i < t1; i++) {
                const TreeNode t = ws.tree[i];
                if (lane == 0) { TileNode o; o.parent = t.parent < 0 ? 0xffffu : (uint16_t)((uint32_t)t.parent - t0); o.len = (uint16_t)t.len; o.node = t.node; tn[i - t0] = o; }
#pragma unroll 1
                for (uint32_t x = lane; x < t.len; x += 32) tb[at + x] = __ldg(ix.seq + t.seq_off + x);
                at += t.len;
            }
#pragma unroll 1
            for (uint32_t x = lane; x < tail_length; x += 32)
                tq[x] = dp_query_base(left_tail ? comp_base(read[tail_length - 1 - x]) : read[e.read_hi + x]);
            if (lane == 0) {
                TileHeader hd; hd.m = tail_length; hd.n_nodes = nt; hd.n_bases = bases; hd.max_gap = g; hd.flags = left_tail ? GB_TILE_LEFT : 0u;
                hd.root_trim = root_trim; hd.bytes = tile_bytes(nt, bases, tail_length); hd.result = ti;
                *reinterpret_cast<TileHeader*>(tile) = hd;
                pp.tile_off[ti] = unit_at;
                TileResult pend; pend.score = 0; pend.status = GB_TILE_ST_PENDING; pend.n_maps = pend.n_edits = pend.path_off = 0; pend.cells_lo = pend.cells_hi = pend.pad = 0;
                pp.results[ti] = pend;
                const int cls = tile_class(tail_length) + (int)wave * TILE_CLASSES;
                const uint32_t pos = atomicAdd(&pp.list_count[cls], 1u);
                if (pos < pp.tile_cap) pp.lists[cls][pos] = ti;
                atomicAdd((unsigned long long*)&pp.stats[3], (unsigned long long)bases * (tail_length + 1));
            }
            unit_at += tile_bytes(nt, bases, tail_length) / 16;
        }
        ti++;
        t0 = t1;
    }
    if (lane == 0) {
        unit_entries[n_entries] = TailPlanEntry{key, first_tile, n_trees, wave};
        atomicAdd((unsigned long long*)&pp.stats[0], 1ull); atomicAdd((unsigned long long*)&pp.stats[1], (unsigned long long)n_trees);
    }
    n_entries++;
    __syncwarp();
    return true;
}

// All plannable tails of one read.
__device__ inline void plan_read(const DevIndex& ix, const MapParamsDev& P, const DevScores& sc, const ReadState& rs, const AlignArgs& a,
                                 const uint8_t* read, uint32_t L, const TailWs& ws, bool paired, uint32_t read_num,
                                 const PlanPools& pp, TailPlanEntry* unit_entries, uint32_t& n_entries) {
    if (rs.status != GB_ITEM_OK || !P.do_dp) return;
    const uint32_t S = rs.item_cnt;
    if (S == 0 || S > MAX_SETS) return;
    const bool deferred = rs.pad[0] > 1;                 // read 2 with tied clusters: align_sets picks a subset of the items later
    int set_score[MAX_SETS]; uint8_t set_order[MAX_SETS];
#pragma unroll 1
    for (uint32_t s = 0; s < S; s++) {
        const uint32_t item = rs.item_off + s;
        if (a.ev.ext_status[item] != GB_ITEM_OK) return;
        set_score[s] = score_extension_group(ev_ext(a.ev, item), a.ev.ext_count[item], L, sc.gap_open, sc.gap_extend);
    }
#pragma unroll 1
    for (uint32_t s = 0; s < S; s++) { uint32_t j = s; while (j > 0 && set_score[s] > set_score[set_order[j - 1]]) { set_order[j] = set_order[j - 1]; j--; } set_order[j] = (uint8_t)s; }
    uint32_t ties = 0;
#pragma unroll 1
    while (ties < S && !(set_score[set_order[0]] > set_score[set_order[ties]])) ties++;
    const double set_cutoff = (double)set_score[set_order[0]] - P.extension_set_score_threshold;
    const uint32_t min_sets = paired ? 2u : (uint32_t)P.min_extension_sets;
    uint32_t unskipped = 0;
#pragma unroll 1
    for (uint32_t oi = 0; oi < S; oi++) {
        const uint32_t s = set_order[oi];
        if (!deferred) {
            bool process;
            if (P.extension_set_score_threshold != 0 && (double)set_score[s] <= set_cutoff) process = unskipped < min_sets;
            else process = unskipped < P.max_alignments || oi < ties;           // any of the tied sets can come first after the shuffle
            if (!process) continue;
            if (!paired && set_score[s] < P.extension_set_min_score) continue;
            unskipped++;
        }
        const uint32_t item = rs.item_off + s;
        const gb_extension* ext = ev_ext(a.ev, item);
        const uint32_t n_ext = a.ev.ext_count[item];
        if (n_ext == 0 || (ext_full(ext[0]) && ext[0].mismatches <= 4)) continue;       // no extensions / direct full-length alignments
        const uint32_t* path_pool = ev_path(a.ev, item);
        uint32_t min_tails = 1;
#pragma unroll 1
        for (uint32_t j = 0; j < n_ext; j++) if (ext_full(ext[j])) min_tails++;
        if (min_tails < 2) min_tails = 2;
        uint8_t eo[64]; const uint32_t ne_ = min(n_ext, 64u);
#pragma unroll 1
        for (uint32_t j = 0; j < ne_; j++) { uint32_t x = j; while (x > 0 && ext[j].score > ext[eo[x - 1]].score) { eo[x] = eo[x - 1]; x--; } eo[x] = (uint8_t)j; }
        const double ecut = (double)ext[eo[0]].score - (double)P.extension_score_threshold;
        const int32_t threshold = ext[eo[0]].score - P.extension_score_threshold;
        uint32_t e_unskipped = 0;
        // The reference skips a partial extension when one has been aligned already, its score is at most `threshold` and its
        // score estimate cannot beat the running winner (:5478-5492).  The extensions above the threshold (tied at the top, in
        // whatever order the LazyRNG puts them) and the first partial one are always aligned: wave 0.  The others wait (wave 1)
        // until tail_decide_kernel has the winner of wave 0 and cancels the ones the reference skips.
        bool partial_aligned = false;
#pragma unroll 1
        for (uint32_t y = 0; y < ne_ && ext[eo[y]].score > threshold; y++) partial_aligned |= !ext_full(ext[eo[y]]);
#pragma unroll 1
        for (uint32_t xi = 0; xi < ne_; xi++) {
            const gb_extension& e = ext[eo[xi]];
            if (P.extension_score_threshold != 0 && (double)e.score <= ecut && e_unskipped >= min_tails) continue;
            e_unskipped++;
            if (ext_full(e)) continue;
            const bool candidate = !deferred && e.score <= threshold && partial_aligned;
#pragma unroll 1
            for (uint32_t side = 0; side < 2; side++) {
                const bool left_tail = side == 0;
                if (e.flags & (left_tail ? GB_EXT_LEFT_FULL : GB_EXT_RIGHT_FULL)) continue;
                plan_tail(ix, P, sc, ws, e, path_pool, read, L, left_tail, tail_key(s, read_num, eo[xi], left_tail), candidate ? 1u : 0u, pp, unit_entries, n_entries);
            }
            if (e.score <= threshold) partial_aligned = true;
        }
    }
}

template <bool PAIRED>
__global__ void __launch_bounds__(ALIGN_WARPS * 32, 4)
tail_plan_kernel(DevIndex ix, MapParamsDev P, DevScores sc, MapBatch b, AlignArgs a, PlanPools pp) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t gwarp = blockIdx.x * ALIGN_WARPS + warp;
    const TailWs ws = carve_tail_ws(a.ws_base + (size_t)gwarp * a.ws_stride, b.Lc, a.tb_cells);
    const uint32_t n_units = PAIRED ? b.n_reads / 2 : b.n_reads;
#pragma unroll 1
    while (true) {
        uint32_t pos = 0, u = 0xffffffffu;
        if (lane == 0) {
            pos = atomicAdd(b.work_counter, 1u);
            if (a.slow_list) u = pos < *a.slow_count ? a.slow_list[pos] : 0xffffffffu; else u = pos;
        }
        pos = __shfl_sync(FULL, pos, 0); u = __shfl_sync(FULL, u, 0);
        if (u >= n_units) break;
        TailPlanEntry* unit_entries = pp.entries + (size_t)pos * PLAN_PER_UNIT;
        uint32_t n_entries = 0;
#pragma unroll 1
        for (uint32_t r = 0; r < (PAIRED ? 2u : 1u); r++) {
            const uint32_t ri = PAIRED ? 2 * u + r : u;
            const ReadState rs = b.states[ri];
            const uint64_t rb = b.read_off[ri];
            const uint32_t L = (uint32_t)(b.read_off[ri + 1] - rb);
            plan_read(ix, P, sc, rs, a, b.reads + rb, L, ws, PAIRED, r, pp, unit_entries, n_entries);
        }
        if (lane == 0) { pp.unit_base[u] = pos * PLAN_PER_UNIT; pp.unit_count[u] = n_entries; }
        __syncwarp();
    }
}


// -----------------------------------------------------------------------------------------------------------------
// tail_decide_kernel: between the two waves of tiles.  For every set of every slow unit it replays
// find_optimal_tail_alignments' bookkeeping (:5369-5622) with the wave-0 results: the running winner after the extensions
// that are always aligned, then the waiting extensions in order — one whose score estimate cannot beat the winner is what
// the reference skips: its tiles are cancelled.  The first waiting extension that survives stops the replay (its own score
// would move the winner), so it and everything behind it run in wave 1.
// -----------------------------------------------------------------------------------------------------------------
__device__ inline void decide_read(const DevIndex& ix, const MapParamsDev& P, const DevScores& sc, const ReadState& rs, const AlignArgs& a,
                                   uint32_t L, uint32_t read_num, const TailPlanEntry* entries, uint32_t n_entries, const uint32_t* tile_off, TileResult* results) {
    if (rs.status != GB_ITEM_OK || rs.pad[0] > 1) return;
    const uint32_t S = rs.item_cnt;
    if (S == 0 || S > MAX_SETS) return;
    auto find_entry = [&](uint32_t key) -> const TailPlanEntry* { for (uint32_t x = 0; x < n_entries; x++) if (entries[x].key == key) return entries + x; return nullptr; };
    // best score over the trees of a finished tail (deterministic_beats only breaks ties): -1 if the tail is not available
    auto tail_score = [&](const TailPlanEntry* pe) -> int32_t {
        int32_t best = 0;
#pragma unroll 1
        for (uint32_t t = 0; t < pe->n_trees; t++) {
            const uint32_t ti = pe->first_tile + t;
            if (tile_off[ti] == TILE_REFUSED) continue;
            if (results[ti].status != GB_TILE_ST_OK) return -1;
            best = max(best, results[ti].score);
        }
        return best;
    };
#pragma unroll 1
    for (uint32_t s = 0; s < S; s++) {
        const uint32_t item = rs.item_off + s;
        if (a.ev.ext_status[item] != GB_ITEM_OK) continue;
        const gb_extension* ext = ev_ext(a.ev, item);
        const uint32_t n_ext = a.ev.ext_count[item];
        if (n_ext == 0 || (ext_full(ext[0]) && ext[0].mismatches <= 4)) continue;
        bool any_waiting = false;
#pragma unroll 1
        for (uint32_t x = 0; x < n_entries; x++) any_waiting |= entries[x].wave == 1 && (entries[x].key >> 9) == ((read_num << 21) | s);
        if (!any_waiting) continue;
        const uint32_t* mism_pool = ev_mism(a.ev, item);
        uint32_t min_tails = 1;
#pragma unroll 1
        for (uint32_t j = 0; j < n_ext; j++) if (ext_full(ext[j])) min_tails++;
        if (min_tails < 2) min_tails = 2;
        Pareto lf[136], rf[136]; uint32_t nl = 0, nr = 0;
#pragma unroll 1
        for (uint32_t j = 0; j < n_ext && nl + 3 < 136; j++) {
            const gb_extension& e = ext[j];
            if (ext_full(e)) continue;
            const int32_t left_penalty = gap_penalty1(e.read_lo, sc);
            const int32_t mid_penalty = (int32_t)e.mism_len * (sc.match + sc.mismatch);
            const int32_t right_penalty = gap_penalty1(L - e.read_hi, sc);
            lf[nl++] = Pareto{e.read_hi, mid_penalty + left_penalty};
            rf[nr++] = Pareto{L - e.read_lo, mid_penalty + right_penalty};
            if (e.mism_len > 0) {
                lf[nl++] = Pareto{mism_pool[e.mism_off], left_penalty};
                rf[nr++] = Pareto{L - mism_pool[e.mism_off + e.mism_len - 1] - 1, right_penalty};
            }
        }
        lf[nl++] = Pareto{ix.k + ix.w - 2, 0}; rf[nr++] = Pareto{ix.k + ix.w - 2, 0};
        nl = find_pareto_frontier(lf, nl); nr = find_pareto_frontier(rf, nr);
        uint8_t eo[64]; const uint32_t ne_ = min(n_ext, 64u);
#pragma unroll 1
        for (uint32_t j = 0; j < ne_; j++) { uint32_t x = j; while (x > 0 && ext[j].score > ext[eo[x - 1]].score) { eo[x] = eo[x - 1]; x--; } eo[x] = (uint8_t)j; }
        const double ecut = (double)ext[eo[0]].score - (double)P.extension_score_threshold;
        uint32_t e_unskipped = 0;
        int32_t winning_score = 0; bool known = true, stop = false;
#pragma unroll 1
        for (uint32_t xi = 0; xi < ne_ && !stop; xi++) {
            const gb_extension& e = ext[eo[xi]];
            if (P.extension_score_threshold != 0 && (double)e.score <= ecut && e_unskipped >= min_tails) continue;
            e_unskipped++;
            const TailPlanEntry* pl = (e.flags & GB_EXT_LEFT_FULL) ? nullptr : find_entry(tail_key(s, read_num, eo[xi], true));
            const TailPlanEntry* pr = (e.flags & GB_EXT_RIGHT_FULL) ? nullptr : find_entry(tail_key(s, read_num, eo[xi], false));
            const bool waiting = (pl && pl->wave == 1) || (pr && pr->wave == 1);
            if (!waiting) {
                // always aligned: its total moves the winner (a full-length extension: its own score)
                int32_t total = e.score;
                if (!ext_full(e)) {
                    if ((!(e.flags & GB_EXT_LEFT_FULL) && !pl) || (!(e.flags & GB_EXT_RIGHT_FULL) && !pr)) { known = false; break; }     // aligned in place: winner unknown here
                    const int32_t ls = pl ? tail_score(pl) : 0, rsx = pr ? tail_score(pr) : 0;
                    if (ls < 0 || rsx < 0) { known = false; break; }
                    total += ls + rsx;
                }
                winning_score = max(winning_score, total);
                continue;
            }
            if (!known) break;
            int32_t estimate = (int32_t)L * sc.match + 2 * sc.full_length_bonus - (int32_t)e.mism_len * (sc.match + sc.mismatch);
            if (!(e.flags & GB_EXT_LEFT_FULL)) estimate -= flank_penalty(e.read_lo, lf, nl, sc);
            if (!(e.flags & GB_EXT_RIGHT_FULL)) estimate -= flank_penalty(L - e.read_hi, rf, nr, sc);
            if (estimate <= winning_score) {
                if (lane_id() == 0) {
#pragma unroll 1
                    for (const TailPlanEntry* pe : {pl, pr}) if (pe) for (uint32_t t = 0; t < pe->n_trees; t++) if (tile_off[pe->first_tile + t] != TILE_REFUSED) results[pe->first_tile + t].status = GB_TILE_ST_CANCELLED;
                }
            } else stop = true;           // this one will be aligned and may move the winner: everything behind it runs
        }
    }
}

template <bool PAIRED>
__global__ void __launch_bounds__(ALIGN_WARPS * 32, 4)
tail_decide_kernel(DevIndex ix, MapParamsDev P, DevScores sc, MapBatch b, AlignArgs a, PlanPools pp) {
    const int lane = threadIdx.x & 31;
    const uint32_t n_units = PAIRED ? b.n_reads / 2 : b.n_reads;
#pragma unroll 1
    while (true) {
        uint32_t pos = 0, u = 0xffffffffu;
        if (lane == 0) {
            pos = atomicAdd(b.work_counter, 1u);
            if (a.slow_list) u = pos < *a.slow_count ? a.slow_list[pos] : 0xffffffffu; else u = pos;
        }
        pos = __shfl_sync(FULL, pos, 0); u = __shfl_sync(FULL, u, 0);
        if (u >= n_units) break;
        const TailPlanEntry* entries = pp.entries + (size_t)pos * PLAN_PER_UNIT;
        const uint32_t n_entries = pp.unit_count[u];
        bool any = false;
#pragma unroll 1
        for (uint32_t x = 0; x < n_entries; x++) any |= entries[x].wave == 1;
        if (!any) continue;
#pragma unroll 1
        for (uint32_t r = 0; r < (PAIRED ? 2u : 1u); r++) {
            const uint32_t ri = PAIRED ? 2 * u + r : u;
            const ReadState rs = b.states[ri];
            const uint32_t L = (uint32_t)(b.read_off[ri + 1] - b.read_off[ri]);
            decide_read(ix, P, sc, rs, a, L, r, entries, n_entries, pp.tile_off, pp.results);
        }
        __syncwarp();
    }
}
