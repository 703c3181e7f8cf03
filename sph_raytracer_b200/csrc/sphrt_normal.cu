// sphrt_normal.cu -- C ABI entry of the fused normal-equations step (its own translation unit:
// the kernels of the walkers are instantiated per operation, see sphrt_forward.cu).
#include "sphrt_kernels.cuh"

extern "C" {

int sphrt_normal(const sphrt_grid* grid, const sphrt_rays* rays, const void* density, int dtype,
                 int64_t n_vol, int64_t rays_per_vol, void* out, const void* b, const void* mask,
                 double scale, void* grad, void* scratch, int64_t scratch_bytes, void* stream)
{
    int rc = check_grid(grid); if (rc) return rc;
    rc = check_rays(rays); if (rc) return rc;
    if (n_vol < 1 || rays_per_vol < 0 || (n_vol > 1 && rays_per_vol == 0)) return SPHRT_EINVAL;
    if (rays->n_rays > 0 && (!density || !out || !grad)) return SPHRT_EINVAL;
    if (dtype == SPHRT_F32)
        return launch_normal<float>(grid, rays, (const float*)density, n_vol, rays_per_vol, (float*)out,
                                    (const float*)b, (const float*)mask, scale, (float*)grad,
                                    scratch, scratch_bytes, stream);
    if (dtype == SPHRT_F64)
        return launch_normal<double>(grid, rays, (const double*)density, n_vol, rays_per_vol, (double*)out,
                                     (const double*)b, (const double*)mask, scale, (double*)grad,
                                     scratch, scratch_bytes, stream);
    return SPHRT_EINVAL;
}

}  // extern "C"
