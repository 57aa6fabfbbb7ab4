// compact.cuh — pack the per-read padded output pools of the align kernels into dense pools so
// only the bytes a GAM record needs cross PCIe (32 B header + 8 B per mapping + 4 B per edit).
// (included by map.cu inside namespace gb)
#pragma once
// <cub/cub.cuh> is included by map.cu at global scope (this header lives inside namespace gb)

struct CountMappings { const gb_alignment* a; __host__ __device__ uint64_t operator()(uint32_t i) const { return a[i].n_mappings; } };
struct CountEdits { const gb_alignment* a; __host__ __device__ uint64_t operator()(uint32_t i) const { return a[i].n_edits; } };

// One warp per read: copy its mappings / edits to their scanned offsets and patch the header.
// n_reads = records (reads x max_multimaps, rank-major: record j * per_rank + read), per_rank = reads of the chunk.
__global__ void compact_gather_kernel(uint32_t n_reads, uint32_t per_rank, gb_alignment* aln, const gb_mapping* maps, const uint32_t* edits,
                                      uint32_t map_cap, uint32_t edit_cap, const uint64_t* map_off, const uint64_t* edit_off,
                                      gb_mapping* out_maps, uint64_t out_map_cap, uint32_t* out_edits, uint64_t out_edit_cap,
                                      const uint64_t* run_base, uint32_t read_base, uint64_t* totals, uint8_t* status) {
    const uint32_t warps_per_block = blockDim.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    // headers carry offsets into the caller's whole pool: run_base = mappings / edits of earlier chunks
    const uint64_t map_base = run_base ? run_base[0] : 0, edit_base = run_base ? run_base[1] : 0;
    for (uint32_t r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < n_reads; r += gridDim.x * warps_per_block) {
        gb_alignment a = aln[r];
        const uint64_t mo = map_off[r], eo = edit_off[r];
        const bool fits = mo + a.n_mappings <= out_map_cap && eo + a.n_edits <= out_edit_cap;
        if (fits) {
            const gb_mapping* sm = maps + (size_t)r * map_cap; const uint32_t* se = edits + (size_t)r * edit_cap;
            for (uint32_t i = lane; i < a.n_mappings; i += 32) out_maps[mo + i] = sm[i];
            for (uint32_t i = lane; i < a.n_edits; i += 32) out_edits[eo + i] = se[i];
        }
        if (lane == 0) {
            a.mapping_off = (uint32_t)(map_base + mo); a.edit_off = (uint32_t)(edit_base + eo); a.read_id = read_base + r % per_rank;
            if (!fits) { a.n_mappings = 0; a.n_edits = 0; a.flags &= ~GB_ALN_MAPPED; a.score = 0; a.mapq = 0; status[r % per_rank] = GB_ITEM_OUT_FULL; }
            aln[r] = a;
            if (r == n_reads - 1) { totals[0] = mo + a.n_mappings; totals[1] = eo + a.n_edits; }
        }
    }
}

// max_multimaps > 1: every record of rank >= 1 starts out absent; the align kernels overwrite the ones that exist.
__global__ void init_absent_kernel(gb_alignment* aln, uint32_t per_rank, uint32_t n_records, uint32_t map_cap, uint32_t edit_cap) {
    for (uint32_t R = per_rank + blockIdx.x * blockDim.x + threadIdx.x; R < n_records; R += gridDim.x * blockDim.x) {
        gb_alignment a;
        a.read_id = R % per_rank; a.score = 0; a.mapq = 0; a.flags = GB_ALN_ABSENT; a.n_mappings = 0; a.mapping_off = R * map_cap; a.edit_off = R * edit_cap;
        a.n_edits = 0; a.mapq_uncapped = 0.f; a.mapq_explored_cap = 0.f;
        aln[R] = a;
    }
}

// After a chunk: run += totals (host-buffer calls keep the running totals on the device so chunks
// can be queued without a host round trip).
__global__ void advance_run_kernel(uint64_t* run, const uint64_t* totals) { run[0] += totals[0]; run[1] += totals[1]; }
__global__ void rebase_offsets_kernel(uint64_t* off, uint32_t n, uint64_t b0) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) off[i] -= b0;
}

// Scans + gather on the handle's stream.  d_totals[0..1] receive the mappings / edits used by
// this chunk.  Temporary storage comes from the handle.
inline int compact_outputs(gb_device* d, uint32_t n_reads, uint32_t per_rank, gb_alignment* d_aln, const gb_mapping* d_maps, const uint32_t* d_edits,
                           uint32_t map_cap, uint32_t edit_cap, gb_mapping* out_maps, uint64_t out_map_cap, uint32_t* out_edits,
                           uint64_t out_edit_cap, const uint64_t* d_run_base, uint32_t read_base, uint64_t* d_totals, uint8_t* d_status) {
    int rc;
    if ((rc = d->c_map_off.reserve(n_reads))) return rc;
    if ((rc = d->c_edit_off.reserve(n_reads))) return rc;
    cub::CountingInputIterator<uint32_t> counting(0);
    cub::TransformInputIterator<uint64_t, CountMappings, cub::CountingInputIterator<uint32_t>> it_m(counting, CountMappings{d_aln});
    cub::TransformInputIterator<uint64_t, CountEdits, cub::CountingInputIterator<uint32_t>> it_e(counting, CountEdits{d_aln});
    size_t tmp_bytes = 0;
    GB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, it_m, d->c_map_off.ptr, (int)n_reads, d->stream));
    if ((rc = d->c_tmp.reserve(tmp_bytes + 256))) return rc;
    size_t tb = d->c_tmp.cap;
    GB_CUDA(cub::DeviceScan::ExclusiveSum(d->c_tmp.ptr, tb, it_m, d->c_map_off.ptr, (int)n_reads, d->stream));
    tb = d->c_tmp.cap;
    GB_CUDA(cub::DeviceScan::ExclusiveSum(d->c_tmp.ptr, tb, it_e, d->c_edit_off.ptr, (int)n_reads, d->stream));
    d->launches += 2;
    const uint32_t grid = std::min<uint32_t>((uint32_t)d->n_sms * 8, (n_reads + 7) / 8);
    compact_gather_kernel<<<grid ? grid : 1, 256, 0, d->stream>>>(n_reads, per_rank, d_aln, d_maps, d_edits, map_cap, edit_cap, d->c_map_off.ptr,
                                                               d->c_edit_off.ptr, out_maps, out_map_cap, out_edits, out_edit_cap,
                                                               d_run_base, read_base, d_totals, d_status);
    d->launches++;
    GB_CUDA(cudaGetLastError());
    return GB_OK;
}
