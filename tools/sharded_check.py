"""ShardedOperator on real GPUs (torchrun, NCCL): static and dynamic volumes against the
single-GPU operator evaluated on every rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \\
        --master-port 29511 tools/sharded_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sph_raytracer_b200 as S                                      # noqa: E402
from sph_raytracer_b200 import parallel as P                        # noqa: E402


def camera(k, n_obs, npix):
    ang = 2 * torch.pi * k / n_obs
    pos = (4 * float(torch.cos(torch.tensor(ang))), 4 * float(torch.sin(torch.tensor(ang))), 0.7 * (k % 3 - 1))
    return S.ConeRectGeom((npix, npix), pos=pos, fov=(30, 30))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    dev = "cuda"
    n_obs, npix = 8, 96
    geoms = S.ViewGeomCollection(*[camera(k, n_obs, npix) for k in range(n_obs)])
    torch.manual_seed(5)
    worst = 0.0
    for name, grid in (("static", S.SphericalGrid((40, 36, 72), size_r=(0.5, 3))),
                       ("dynamic", S.SphericalGrid((n_obs, 40, 36, 72), size_r=(0.5, 3)))):
        for dtype, tol in ((torch.float32, 2e-5), (torch.float64, 1e-11)):
            x = torch.rand(tuple(grid.shape), dtype=dtype, device=dev)
            dist.broadcast(x, 0)
            w = torch.rand((n_obs, npix, npix), dtype=dtype, device=dev)
            dist.broadcast(w, 0)
            full = S.Operator(grid, geoms, device=dev)
            xr = x.clone().requires_grad_(True)
            yr = full(xr)
            (yr * w).sum().backward()
            sop = P.ShardedOperator(grid, geoms, S.Operator, device=dev)
            xs = x.clone().requires_grad_(True)
            ys = sop(xs)
            (ys * w[sop.lo:sop.hi]).sum().backward()
            ty = sop.T(w[sop.lo:sop.hi])
            scale_y, scale_g = float(yr.abs().max()), float(xr.grad.abs().max())
            ey = float((ys - yr[sop.lo:sop.hi]).abs().max()) / scale_y
            eg = float((xs.grad - xr.grad).abs().max()) / scale_g
            et = float((ty - xr.grad).abs().max()) / scale_g
            errs = torch.tensor([ey, eg, et], device=dev, dtype=torch.float64)
            dist.all_reduce(errs, op=dist.ReduceOp.MAX)
            ok = bool((errs <= tol).all())
            worst = max(worst, float(errs.max()) / tol)
            if rank == 0:
                print(f"{name:8s} {str(dtype):14s} world {world}  obs/rank {sop.hi - sop.lo}  walker {full.walker()}"
                      f"  rel err forward {errs[0]:.2e}  gradient {errs[1]:.2e}  T {errs[2]:.2e}  tol {tol:g}  "
                      f"{'ok' if ok else 'FAIL'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if worst <= 1 else 1)


if __name__ == "__main__":
    main()
