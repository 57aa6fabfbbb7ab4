// device_state.cuh — the opaque gb_device handle: HBM-resident index, stream, workspaces.
#pragma once
#include "giraffe_b200.h"
#include "device_index.cuh"
#include "extend.cuh"
#include "map_state.cuh"
#include "xdrop_tile.cuh"
namespace gb { struct DbgCluster; }

#include <cuda_runtime.h>
#include <string>
#include <vector>

namespace gb {

constexpr int EXTEND_WARPS = 8;           // warps per CTA of the extension kernel
constexpr uint32_t EXTEND_Q_CAP = 512;    // frontier entries per warp (64 B each)
constexpr uint32_t EXTEND_A_CAP = 4096;   // path-arena nodes per warp (8 B each)

extern thread_local std::string g_last_error;
int fail_cuda(cudaError_t e, const char* what);

#define GB_CUDA(call)                                                        \
    do {                                                                     \
        cudaError_t _e = (call);                                             \
        if (_e != cudaSuccess) return ::gb::fail_cuda(_e, #call);            \
    } while (0)

// Growable device buffer (cudaMalloc on demand, never shrinks).
template <typename T>
struct DevBuf {
    T* ptr = nullptr;
    size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() { if (ptr) cudaFree(ptr); ptr = nullptr; cap = 0; }
    int reserve(size_t n) {
        if (n <= cap && ptr) return GB_OK;
        if (ptr) { cudaFree(ptr); ptr = nullptr; cap = 0; }
        if (n == 0) n = 1;
        cudaError_t e = cudaMalloc(&ptr, n * sizeof(T));
        if (e != cudaSuccess) { ptr = nullptr; return fail_cuda(e, "cudaMalloc"); }
        cap = n;
        return GB_OK;
    }
    // allocate `n` elements, copy `n_copy` (default n) from host
    int upload(const T* host, size_t n, cudaStream_t s, size_t n_copy = (size_t)-1) {
        int rc = reserve(n);
        if (rc) return rc;
        if (n_copy == (size_t)-1) n_copy = n;
        if (n_copy && host) {
            cudaError_t e = cudaMemcpyAsync(ptr, host, n_copy * sizeof(T), cudaMemcpyHostToDevice, s);
            if (e != cudaSuccess) return fail_cuda(e, "cudaMemcpyAsync H2D");
        }
        return GB_OK;
    }
};

} // namespace gb

struct gb_device {
    int device = 0;
    int n_sms = 0;
    std::vector<uint32_t> h_node_len;      // host copy of the node lengths (workspace sizing of the DP seams)
    std::vector<gb_dist_payload> h_dist;   // host copy of the distance payload (fragment-length training)
    // first-pass seeding table sizes (minimizers, clusters per read); GIRAFFE_B200_SEED_TABLES="Mc,Cc" overrides
    uint32_t seed_mc = 64, seed_cc = 16, seed_ns = 64;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    cudaEvent_t ev_stage[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    gb::DevIndex ix{};
    gb::DevScores sc{1, 4, 6, 1, 5};
    gb::DevBuf<gb_node_rec> nodes;
    gb::DevBuf<uint8_t> seq;
    gb::DevBuf<uint32_t> gbwt;
    gb::DevBuf<gb_dist_payload> dist;
    gb::DevBuf<gb_min_cell> table;
    gb::DevBuf<gb_hit> hits;
    gb::DevBuf<uint32_t> slot_order;
    gb::DevBuf<gb_slot_rec> slots; gb::DevBuf<uint16_t> site_dist;
    std::vector<gb_slot_rec> h_slots; std::vector<uint16_t> h_site_dist;      // host copies (fragment-length training)
    gb::DevBuf<gb::QEntry> ws_queue;
    gb::DevBuf<gb::ArenaNode> ws_arena;
    gb::DevBuf<uint32_t> p_post;                     // items extend_kernel hands to extend_finish_kernel
    uint32_t* work_counter = nullptr;
    float last_kernel_ms = 0.f;
    uint64_t launches = 0;
    // mapping pipeline: parameter tables, intermediate pools, workspaces, I/O staging
    bool tables_ready = false; uint32_t tables_hard_hit_cap = 0;
    gb::DevBuf<double> t_hit, t_plo, t_phred;
    gb::DevBuf<gb::ReadState> p_states;
    gb::DevBuf<gb::DevMinimizer> p_min;
    gb::DevBuf<gb::DevSeed> p_seeds;
    gb::DevBuf<gb::DevItem> p_items;
    gb::DevBuf<gb_seed> p_ext_seeds;
    gb::DevBuf<uint32_t> p_cursors, p_ext_count, p_path, p_mism;
    gb::DevBuf<uint8_t> p_ext_status;
    gb::DevBuf<uint32_t> p_big_list, p_big_of, p_big_path, p_big_mism; gb::DevBuf<gb_extension> p_big_ext;      // ExtendBig
    gb::DevBuf<gb_extension> p_ext;
    gb::DevBuf<uint8_t> ws_tail, ws_cand, w_reads, w_quals, ws_rescue;
    gb::DevBuf<gb::PairState> p_pairs;
    gb::DevBuf<uint32_t> p_retry;         // units the first seeding pass could not fit
    gb::DevBuf<uint32_t> p_slow;          // pairs routed to the warp-per-pair align kernel
    gb::DevBuf<uint32_t> p_rescue;        // pairs the plain kernel defers to the rescue instantiation
    gb::DevBuf<gb_mapping> pad_maps;
    gb::DevBuf<uint32_t> pad_edits;
    gb::DevBuf<uint64_t> c_map_off, c_edit_off, c_totals;
    gb::DevBuf<uint8_t> c_tmp;
    // host-buffer entry points: two staging sets so chunk i+1 uploads and chunk i-1 downloads while
    // chunk i computes (streams s_in / stream / s_out)
    struct IoSet {
        gb::DevBuf<uint8_t> reads, quals, status;
        gb::DevBuf<uint64_t> read_off, totals;
        gb::DevBuf<gb_alignment> aln;
        gb::DevBuf<gb_mapping> maps;
        gb::DevBuf<uint32_t> edits;
        cudaEvent_t ev_in = nullptr, ev_done = nullptr, ev_hdr = nullptr, ev_out = nullptr, ev_k0 = nullptr, ev_k1 = nullptr;
        void release() { reads.release(); quals.release(); status.release(); read_off.release(); totals.release(); aln.release(); maps.release(); edits.release(); }
    } io[2];
    cudaStream_t s_in = nullptr, s_out = nullptr;
    uint64_t* h_totals = nullptr;          // pinned, 3 per set: mappings, edits, pool-overflow flag
    double pool_scale = 1.0;               // intermediate pools = per-read averages x this; doubled when a chunk overflows (GIRAFFE_B200_POOL_SCALE)
    uint32_t pool_reruns = 0;              // chunks redone because of it
    gb::DbgCluster* dbg_clusters = nullptr; bool debug_stop_after_seed = false;     // gb_debug_seed_stage
    // tail plan + DP tiles (xdrop_tile.cuh); GIRAFFE_B200_TILES=0 keeps every tail DP in the align kernels (int32 sweep)
    bool use_tiles = true; uint32_t last_tile_problems = 0;
    int extend_minb = 4, fast_minb = 12;          // launch-bounds instantiations (GIRAFFE_B200_EXTEND_MINB / _FAST_MINB)
    gb_alignment* mirror_aln = nullptr; gb_mapping* mirror_maps = nullptr; uint32_t* mirror_edits = nullptr; uint64_t mirror_map_cap = 0, mirror_edit_cap = 0;
    gb::DevBuf<gb::TailPlanEntry> pl_entries;
    gb::DevBuf<uint32_t> pl_unit_base, pl_unit_count, pl_tile_off, pl_lists, pl_paths;
    gb::DevBuf<uint8_t> pl_tiles, ws_tile;
    gb::DevBuf<gb::TileResult> pl_results;
    gb::DevBuf<uint64_t> pl_stats;
    // per-kernel device times of the last map_device call (events between the launches)
    static constexpr int KT_MAX = 48;
    cudaEvent_t kt_ev[KT_MAX] = {}; const char* kt_name[KT_MAX] = {}; int kt_n = 0;
    void kt_reset() { kt_n = 0; }
    int kt_mark(const char* name) {
        if (kt_n >= KT_MAX) return GB_OK;
        if (!kt_ev[kt_n]) { cudaError_t e = cudaEventCreate(&kt_ev[kt_n]); if (e != cudaSuccess) return gb::fail_cuda(e, "cudaEventCreate"); }
        cudaError_t e = cudaEventRecord(kt_ev[kt_n], stream);
        if (e != cudaSuccess) return gb::fail_cuda(e, "cudaEventRecord");
        kt_name[kt_n++] = name;
        return GB_OK;
    }
    gb::DevBuf<uint64_t> c_run;            // running mapping / edit totals of a host-buffer call
    uint32_t map_chunk = 1u << 20;         // reads per chunk; GIRAFFE_B200_MAP_CHUNK overrides
    void release_all() {
        nodes.release(); seq.release(); gbwt.release(); dist.release(); table.release(); hits.release(); slot_order.release(); slots.release(); site_dist.release();
        ws_queue.release(); ws_arena.release(); p_post.release();
        t_hit.release(); t_plo.release(); t_phred.release(); p_states.release(); p_min.release(); p_seeds.release();
        p_items.release(); p_ext_seeds.release(); p_cursors.release(); p_ext_count.release(); p_path.release(); p_mism.release();
        p_ext_status.release(); p_ext.release(); p_big_list.release(); p_big_of.release(); p_big_path.release(); p_big_mism.release(); p_big_ext.release(); ws_tail.release(); ws_cand.release(); ws_rescue.release(); w_reads.release(); w_quals.release(); p_pairs.release(); p_slow.release(); p_rescue.release(); p_retry.release();
        pad_maps.release(); pad_edits.release(); c_map_off.release(); c_edit_off.release(); c_totals.release(); c_tmp.release();
        io[0].release(); io[1].release(); c_run.release();
        pl_entries.release(); pl_unit_base.release(); pl_unit_count.release(); pl_tile_off.release(); pl_lists.release(); pl_paths.release(); pl_tiles.release(); ws_tile.release(); pl_results.release(); pl_stats.release();
    }
};
