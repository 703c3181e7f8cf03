// sphrt_adjoint.cu -- C ABI entry of the adjoint pass (its kernels are instantiated here).
#include "sphrt_kernels.cuh"

extern "C" {

int sphrt_adjoint(const sphrt_grid* grid, const sphrt_rays* rays, const void* y, int dtype,
                  int64_t n_vol, int64_t rays_per_vol, void* grad, void* scratch,
                  int64_t scratch_bytes, void* stream)
{
    int rc = check_grid(grid); if (rc) return rc;
    rc = check_rays(rays); if (rc) return rc;
    if (n_vol < 1 || rays_per_vol < 0) return SPHRT_EINVAL;
    if (rays->n_rays > 0 && (!y || !grad)) return SPHRT_EINVAL;
    if (dtype == SPHRT_F32)
        return launch_volume_op<AdjointOp, float>(grid, rays, (const float*)y, (float*)grad, n_vol,
                                                  rays_per_vol, false, scratch, scratch_bytes, stream);
    if (dtype == SPHRT_F64)
        return launch_volume_op<AdjointOp, double>(grid, rays, (const double*)y, (double*)grad, n_vol,
                                                   rays_per_vol, false, scratch, scratch_bytes, stream);
    return SPHRT_EINVAL;
}

}  // extern "C"
