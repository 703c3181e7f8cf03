// sphrt_forward.cu -- C ABI entry of the forward pass (its kernels are instantiated here).
#include "sphrt_kernels.cuh"

extern "C" {

int sphrt_forward(const sphrt_grid* grid, const sphrt_rays* rays, const void* density, int dtype,
                  int64_t n_vol, int64_t rays_per_vol, void* out, void* scratch,
                  int64_t scratch_bytes, void* stream)
{
    int rc = check_grid(grid); if (rc) return rc;
    rc = check_rays(rays); if (rc) return rc;
    if (n_vol < 1 || rays_per_vol < 0) return SPHRT_EINVAL;
    if (rays->n_rays > 0 && (!density || !out)) return SPHRT_EINVAL;
    if (dtype == SPHRT_F32)
        return launch_volume_op<ForwardOp, float>(grid, rays, (const float*)density, (float*)out, n_vol,
                                                  rays_per_vol, true, scratch, scratch_bytes, stream);
    if (dtype == SPHRT_F64)
        return launch_volume_op<ForwardOp, double>(grid, rays, (const double*)density, (double*)out, n_vol,
                                                   rays_per_vol, true, scratch, scratch_bytes, stream);
    return SPHRT_EINVAL;
}

}  // extern "C"
