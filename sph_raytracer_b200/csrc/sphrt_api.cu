// sphrt_api.cu -- the rest of the C ABI of libsphrt.so: queries, counting / trace / store /
// event exports, stored-operator passes, find_starts.
#include "sphrt_kernels.cuh"

int sphrt_g_last_cuda_error = 0;
unsigned long long sphrt_g_launches = 0;

// ---- C ABI --------------------------------------------------------------------------------------

extern "C" {

int sphrt_version(void) { return SPHRT_VERSION; }

int sphrt_last_cuda_error(void) { return sphrt_g_last_cuda_error; }

unsigned long long sphrt_launch_count(void) { return sphrt_g_launches; }

const char* sphrt_error_string(int code)
{
    switch (code) {
    case SPHRT_OK: return "ok";
    case SPHRT_EINVAL: return "invalid argument";
    case SPHRT_ETOOBIG: return "grid axis or ray count too large";
    case SPHRT_ESCRATCH: return "scratch buffer missing or smaller than sphrt_scratch_bytes()";
    case SPHRT_ECUDA: return "CUDA launch failed";
    default: return "unknown error";
    }
}

int64_t sphrt_scratch_bytes(const sphrt_grid* grid, int64_t n_rays)
{
    if (grid == nullptr || grid->ne < 1 || n_rays < 0) return 0;
    // masks + counters + global cone lists (very large grids) + records of the warp walker
    return scratch_bytes_before_recs(grid, n_rays) + warp_scratch_bytes(*grid, n_rays);
}



int64_t sphrt_plan_cache_bytes(const sphrt_grid* grid, int64_t n_rays)
{
    if (check_grid(grid) != SPHRT_OK || n_rays <= 0) return 0;
    if (!plan_warp(*grid).ok || !plan_coop(*grid, n_rays).ok) return 0;
    return plan_cache_layout(n_rays).bytes;
}

int sphrt_count_segments(const sphrt_grid* grid, const sphrt_rays* rays, int32_t* count,
                         unsigned long long* total, void* scratch, int64_t scratch_bytes, void* stream)
{
    int rc = check_grid(grid); if (rc) return rc;
    rc = check_rays(rays); if (rc) return rc;
    if (rays->n_rays > 0 && !count) return SPHRT_EINVAL;
    const CountOp op{count, total};
    return launch_passes(grid, rays, op, scratch, scratch_bytes, stream);
}

int sphrt_trace_segments(const sphrt_grid* grid, const sphrt_rays* rays, int64_t max_seg,
                         int32_t* vox, double* seg_len, int32_t* count, void* scratch,
                         int64_t scratch_bytes, void* stream)
{
    int rc = check_grid(grid); if (rc) return rc;
    rc = check_rays(rays); if (rc) return rc;
    if (max_seg < 1) return SPHRT_EINVAL;
    if (rays->n_rays > 0 && (!vox || !seg_len || !count)) return SPHRT_EINVAL;
    const TraceOp op{vox, seg_len, count, max_seg};
    return launch_passes(grid, rays, op, scratch, scratch_bytes, stream);
}

int sphrt_store_segments(const sphrt_grid* grid, const sphrt_rays* rays, const int64_t* offsets,
                         int32_t* vox, void* seg_len, int len_dtype, void* scratch,
                         int64_t scratch_bytes, void* stream)
{
    int rc = check_grid(grid); if (rc) return rc;
    rc = check_rays(rays); if (rc) return rc;
    if (rays->n_rays > 0 && (!offsets || !vox || !seg_len)) return SPHRT_EINVAL;
    if (len_dtype == SPHRT_F32) {
        const StoreOp<float> op{offsets, vox, (float*)seg_len, grid->ne, grid->na};
        return launch_passes(grid, rays, op, scratch, scratch_bytes, stream);
    }
    if (len_dtype == SPHRT_F64) {
        const StoreOp<double> op{offsets, vox, (double*)seg_len, grid->ne, grid->na};
        return launch_passes(grid, rays, op, scratch, scratch_bytes, stream);
    }
    return SPHRT_EINVAL;
}

int sphrt_forward_stored(const int64_t* offsets, const int32_t* vox, const void* seg_len, int len_dtype,
                         int64_t n_rays, const void* density, int dtype, int64_t n_vol,
                         int64_t rays_per_vol, int64_t nvox, void* out, void* stream)
{
    return dispatch_stored<false>(offsets, vox, seg_len, len_dtype, n_rays, density, dtype, out, n_vol,
                                  rays_per_vol, nvox, stream);
}

int sphrt_adjoint_stored(const int64_t* offsets, const int32_t* vox, const void* seg_len, int len_dtype,
                         int64_t n_rays, const void* y, int dtype, int64_t n_vol,
                         int64_t rays_per_vol, int64_t nvox, void* grad, void* stream)
{
    return dispatch_stored<true>(offsets, vox, seg_len, len_dtype, n_rays, y, dtype, grad, n_vol,
                                 rays_per_vol, nvox, stream);
}

int sphrt_pack_count(const sphrt_grid* grid, const sphrt_rays* rays, int32_t* count,
                     unsigned long long* total, void* scratch, int64_t scratch_bytes, void* stream)
{
    int rc = check_grid(grid); if (rc) return rc;
    rc = check_rays(rays); if (rc) return rc;
    if ((int64_t)grid->nr * grid->ne * grid->na > (int64_t)kPackLenMask) return SPHRT_ETOOBIG;
    if (rays->n_rays > 0 && !count) return SPHRT_EINVAL;
    const PackOp<false> op{nullptr, nullptr, count, total, grid->ne, grid->na};
    return launch_passes(grid, rays, op, scratch, scratch_bytes, stream);
}

int sphrt_pack_segments(const sphrt_grid* grid, const sphrt_rays* rays, const int64_t* offsets,
                        uint32_t* words, void* scratch, int64_t scratch_bytes, void* stream)
{
    int rc = check_grid(grid); if (rc) return rc;
    rc = check_rays(rays); if (rc) return rc;
    if ((int64_t)grid->nr * grid->ne * grid->na > (int64_t)kPackLenMask) return SPHRT_ETOOBIG;
    if (rays->n_rays > 0 && (!offsets || !words)) return SPHRT_EINVAL;
    const PackOp<true> op{offsets, words, nullptr, nullptr, grid->ne, grid->na};
    return launch_passes(grid, rays, op, scratch, scratch_bytes, stream);
}

int sphrt_forward_packed(const int64_t* offsets, const uint32_t* words, int64_t n_rays, int ne, int na,
                         const void* density, int dtype, int64_t n_vol, int64_t rays_per_vol,
                         int64_t nvox, void* out, void* stream)
{
    return dispatch_packed<false>(offsets, words, n_rays, ne, na, density, dtype, out, n_vol, rays_per_vol, nvox, stream);
}

int sphrt_adjoint_packed(const int64_t* offsets, const uint32_t* words, int64_t n_rays, int ne, int na,
                         const void* y, int dtype, int64_t n_vol, int64_t rays_per_vol,
                         int64_t nvox, void* grad, void* stream)
{
    return dispatch_packed<true>(offsets, words, n_rays, ne, na, y, dtype, grad, n_vol, rays_per_vol, nvox, stream);
}

int sphrt_trace_events(const sphrt_grid* grid, const sphrt_rays* rays, int64_t max_ev,
                       int32_t* vox, double* ev_len, int32_t* count, unsigned long long* stats,
                       void* scratch, int64_t scratch_bytes, void* stream)
{
    int rc = check_grid(grid); if (rc) return rc;
    rc = check_rays(rays); if (rc) return rc;
    if (max_ev < 1) return SPHRT_EINVAL;
    if (rays->n_rays > 0 && (!vox || !ev_len || !count)) return SPHRT_EINVAL;
    const EventTraceOp op{vox, ev_len, count, max_ev, grid->nr};
    return launch_passes(grid, rays, op, scratch, scratch_bytes, stream, true, stats);
}

int sphrt_walker_kind(const sphrt_grid* grid, int64_t n_rays)
{
    if (check_grid(grid) != SPHRT_OK) return SPHRT_EINVAL;
    return (plan_coop(*grid, n_rays).ok && want_coop(*grid)) ? 1 : 0;
}

int sphrt_candidates(const sphrt_grid* grid, const sphrt_rays* rays, double* t, int32_t* region,
                     int32_t* geo, void* stream)
{
    int rc = check_grid(grid); if (rc) return rc;
    rc = check_rays(rays); if (rc) return rc;
    if (rays->n_rays == 0) return SPHRT_OK;
    if (!t || !region || !geo) return SPHRT_EINVAL;
    const KRays kr = make_krays(rays);
    const size_t smem = (size_t)table_doubles(*grid) * sizeof(double);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int64_t blocks = (rays->n_rays + kBlock - 1) / kBlock;
    candidates_kernel<<<(unsigned)blocks, kBlock, smem, (cudaStream_t)stream>>>(*grid, kr, t, region, geo);
    return cuda_status();
}

int sphrt_find_starts(const sphrt_grid* grid, int64_t n, const double* xs, int32_t* out, void* stream)
{
    int rc = check_grid(grid); if (rc) return rc;
    if (n < 0 || (n > 0 && (!xs || !out))) return SPHRT_EINVAL;
    if (n == 0) return SPHRT_OK;
    const int64_t blocks = (n + kBlock - 1) / kBlock;
    const size_t smem = (size_t)table_doubles(*grid) * sizeof(double);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(find_starts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    find_starts_kernel<<<(unsigned)blocks, kBlock, smem, (cudaStream_t)stream>>>(*grid, n, xs, out);
    return cuda_status();
}

int sphrt_export_rays(const sphrt_rays* rays, double* dirs_out, void* stream)
{
    int rc = check_rays(rays); if (rc) return rc;
    if (rays->n_rays == 0) return SPHRT_OK;
    if (!dirs_out) return SPHRT_EINVAL;
    const KRays kr = make_krays(rays);
    const int64_t blocks = (rays->n_rays + 255) / 256;
    export_rays_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(kr, dirs_out);
    return cuda_status();
}

}  // extern "C"
