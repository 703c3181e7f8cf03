#!/usr/bin/env python3
"""Benchmark of the forward + adjoint line-integration hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c2|c1|c4] [--obs M]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference        # CPU oracle port of the reference, bounded sample

One step = one forward pass + one adjoint pass over every ray of the workload (the
retrieval loop's unit of work, reference retrieval.py:90-116).  Metric: ray-segments/s for
forward + adjoint, a ray-segment being one merged non-zero voxel intersection
(SURVEY.md 8d), counted once per pass.  Prints ONE JSON line on rank 0.

Default workload = BASELINE.json configs[4], the one its metric is quoted on at 1/2/4/8
GPUs: grid (200,180,360), r in (3,25); 1000 observations of 512x512 pinhole detectors at
|pos| = 100, fov 30 deg, seeded directions; density / adjoint input rand fp32.  The
observations are sharded over the ranks (strong scaling), the adjoint ends with one NCCL
all-reduce of the gradient volume.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FWD_BYTES = (4, 52, 4)      # per segment, per ray, per voxel   
ADJ_BYTES = (8, 52, 8)


# ---- workloads ------------------------------------------------------------------------------
def workload(name, n_obs=None):
    """(grid, list of per-observation geometries, description dict)"""
    import sph_raytracer_b200 as S
    if name == "c5":
        n_obs = n_obs or 1000
        grid = S.SphericalGrid(shape=(200, 180, 360), size_r=(3, 25))
        gen = torch.Generator().manual_seed(0)
        theta = torch.rand(1000, generator=gen, dtype=torch.float64) * 2 * math.pi
        phi = (0.2 + 0.6 * torch.rand(1000, generator=gen, dtype=torch.float64)) * math.pi
        geoms = [S.ConeRectGeom((512, 512),
                                (100 * math.cos(t) * math.sin(p), 100 * math.sin(t) * math.sin(p),
                                 100 * math.cos(p)), fov=(30, 30))
                 for t, p in zip(theta[:n_obs].tolist(), phi[:n_obs].tolist())]
        desc = {"workload": f"large synthetic sweep: grid (200,180,360) r in (3,25), {n_obs} obs x "
                            "512x512 ConeRectGeom |pos|=100 fov 30 (BASELINE configs[4])"}
    elif name in ("c2", "c4"):
        n_obs = n_obs or (50 if name == "c2" else 20)
        shape = (50, 50, 50) if name == "c2" else (n_obs, 50, 50, 50)
        grid = S.SphericalGrid(shape=shape)
        geoms = [S.ConeCircGeom((100, 50), (5 * torch.cos(t), 5 * torch.sin(t), 1), fov=(0, 45))
                 for t in torch.linspace(0, 2 * torch.pi, n_obs)]
        desc = {"workload": f"{'static_retrieval/memory_usage' if name == 'c2' else 'dynamic_measurements'}"
                            f": grid {shape}, {n_obs} obs x ConeCircGeom (100,50) orbit r=5 z=1 "
                            f"(BASELINE configs[{1 if name == 'c2' else 3}])"}
    elif name == "c1":
        grid = S.SphericalGrid(shape=(50, 50, 50))
        geoms = [S.ConeRectGeom((256, 256), (5, 0, 0), fov=(45, 45))]
        desc = {"workload": "single_vantage: grid (50,50,50), ConeRectGeom (256,256) pos (5,0,0) "
                            "fov 45 (BASELINE configs[0])"}
    else:
        raise SystemExit(f"unknown workload {name}")
    return grid, geoms, desc


# ---- clocks -----------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples nvidia-smi SM clock and throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([f.strip() for f in out.strip().split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        clocks, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            if len(s) < 6:
                continue
            try:
                clocks.append(float(s[0])); mx = float(s[1])
            except ValueError:
                continue
            for n, v in zip(names, s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(clocks)) if clocks else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(clocks)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---- CPU baseline: the oracle port of the reference on a bounded sample ------------------------
def cpu_baseline(name, steps=1):
    """Reference algorithm (oracle/sph_oracle.c) on the host: build the stored operator for a
    bounded sample of the workload, then time forward + adjoint on it."""
    from oracle import oracle as O
    O.build()
    grid, geoms, _ = workload(name, n_obs=1)
    g = geoms[0]
    if name == "c5":
        rays = g.rays[::2, ::2]            # 65 536 rays of one 512x512 observation (2.4 GB of regs / lens)
        sample = "1 observation, every 2nd pixel (65536 rays) of the c5 workload"
    else:
        rays = g.rays
        sample = f"1 observation ({rays.shape[0] * rays.shape[1]} rays) of the {name} workload"
    xs, d, _ = O.broadcast_rays(g.ray_starts, rays)
    og = O.OracleGrid(grid)
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    regs, lens = O.trace(og, xs, d, nthreads=cores)
    t_build = time.perf_counter() - t0
    _, _, cnt = O.canonical(og, regs, lens, max_seg=8)
    nseg = int(cnt.sum())
    rng = np.random.default_rng(0)
    dens = rng.random((og.nr, og.ne, og.na), dtype=np.float32)
    y = rng.random((1, len(xs)), dtype=np.float32)
    best = None
    for _ in range(max(steps, 1)):
        t0 = time.perf_counter()
        O.forward(og, regs, lens, dens, nthreads=cores)
        t1 = time.perf_counter()
        O.adjoint(og, regs, lens, y)
        t2 = time.perf_counter()
        best = (t1 - t0, t2 - t1) if best is None or (t2 - t0) < sum(best) else best
    t_fwd, t_adj = best
    return {"value": 2 * nseg / (t_fwd + t_adj), "unit": "ray-segments/s", "cores": cores,
            "kind": "port", "sample": sample + f"; stored operator built once in {t_build:.2f} s "
            f"({len(xs) / t_build:.0f} rays/s), fwd {t_fwd * 1e3:.1f} ms, adjoint {t_adj * 1e3:.1f} ms",
            "segments": nseg, "rays": len(xs), "ms_per_step": (t_fwd + t_adj) * 1e3,
            "build_s": t_build}


def reference_memory_gb(grid, n_rays):
    """Bytes the reference needs for `n_rays` rays through `grid`: every ray owns
    nvox_ray = 2(nr+1) + 2(ne+1) + (na+1) candidate crossings; the stored operator holds 3 int64
    indices + 1 float64 length for each of them and the start point (regs / lens have nvox_ray + 1
    slots), its estimator counts 6 x 8 B per candidate."""
    nr, ne, na = (int(v) for v in tuple(grid.shape)[-3:])
    nvox_ray = 2 * (nr + 1) + 2 * (ne + 1) + (na + 1)
    return {"stored_operator": 32.0 * (nvox_ray + 1) * n_rays / 1e9,
            "estimator": 48.0 * nvox_ray * n_rays / 1e9, "slots_per_ray": nvox_ray + 1}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port) as its own arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_baseline(args.workload, steps=args.steps + args.warmup)
    _, _, desc = workload(args.workload, args.obs)          # the GPU arm's config; the sample is in cpu_baseline
    line = {"impl": "reference", "metric": "ray-segments/s fwd+adjoint", "value": cb["value"],
            "unit": "ray-segments/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64 geometry / f32 volume", "data": "synthetic",
            "config": desc, "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "ray-segments/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---- the GPU arm ----------------------------------------------------------------------------------
def run_gpu(args):
    import torch.distributed as dist
    import sph_raytracer_b200 as S
    from sph_raytracer_b200.parallel import shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    grid, geoms, desc = workload(args.workload, args.obs)
    n_obs = len(geoms)
    lo, hi = shard_bounds(n_obs, rank, world) if n_obs >= world else (0, n_obs)
    dynamic = grid.dynamic
    mine = geoms[lo:hi]
    geom = S.ViewGeomCollection(*mine) if len(mine) > 1 or dynamic else mine[0]
    torch.cuda.reset_peak_memory_stats(dev)
    t0 = time.perf_counter()
    op = S.Operator(grid, geom, device=dev)
    torch.cuda.synchronize(dev)
    t_build = time.perf_counter() - t0
    n_rays = op._rays.n
    _, n_seg = op.segment_counts()

    full_shape = tuple(grid.shape)
    vshape = ((hi - lo,) + full_shape[1:]) if dynamic else full_shape
    gen = torch.Generator().manual_seed(0)
    x_host = torch.rand(vshape, generator=gen).pin_memory()
    y_host = torch.rand(op._rays.shape, generator=torch.Generator().manual_seed(1)).pin_memory()
    out_host = torch.empty(op._rays.shape).pin_memory()
    grad_host = torch.empty(vshape).pin_memory()
    x = x_host.to(dev)
    y = y_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_device():
        out = op._apply_forward(x)
        g = op._apply_adjoint(y, vshape)
        if world > 1 and not dynamic:
            dist.all_reduce(g)
        return out, g

    def step_e2e():
        op.forward(x_host, out=out_host)
        gd = op._apply_adjoint(y_host.to(dev, non_blocking=True), vshape)
        if world > 1 and not dynamic:
            dist.all_reduce(gd)
        grad_host.copy_(gd, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    fwd_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
               torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev[0].record()
    for k in range(args.steps):
        a, b, c = fwd_ev[k]
        a.record()
        out = op._apply_forward(x)
        b.record()
        g = op._apply_adjoint(y, vshape)
        c.record()
        if world > 1 and not dynamic:
            dist.all_reduce(g)
    ev[1].record()
    barrier()
    elapsed_ms = ev[0].elapsed_time(ev[1])
    t_fwd = float(np.mean([a.elapsed_time(b) for a, b, c in fwd_ev]))
    t_adj = float(np.mean([b.elapsed_time(c) for a, b, c in fwd_ev]))
    clocks = sampler.summary()
    if os.environ.get("SPHRT_BENCH_VERBOSE"):
        print("per-step fwd ms:", [round(a.elapsed_time(b), 3) for a, b, c in fwd_ev],
              "adj ms:", [round(b.elapsed_time(c), 3) for a, b, c in fwd_ev], file=sys.stderr)

    # end-to-end: host buffers in, host buffers out, copies inside the timed region
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    peak_bytes = torch.cuda.max_memory_allocated(dev)

    stats = torch.tensor([elapsed_ms, e2e_ms, t_fwd, t_adj, float(peak_bytes)], dtype=torch.float64, device=dev)
    sums = torch.tensor([float(n_seg), float(n_rays)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    elapsed_ms, e2e_ms, t_fwd, t_adj, peak_bytes = stats.tolist()
    tot_seg, tot_rays = sums.tolist()

    # Extra, not the headline: the same step with the STORED operator (Operator(store_segments=True):
    # the reference's own design, segments traced once and kept as an 8-byte-per-segment list),
    # where that list fits next to everything else -- at config 5 from four GPUs on.
    stored = None
    if os.environ.get("SPHRT_BENCH_STORED", "1") != "0" and not dynamic:
        def agree(flag):                 # every rank takes the same branch: minimum over the ranks
            t = torch.tensor([1.0 if flag else 0.0], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
            return t.item() > 0
        out = g = None
        torch.cuda.empty_cache()
        free = torch.cuda.mem_get_info(dev)[0]
        if agree(8.0 * n_seg + 8.0 * n_rays < 0.6 * free):
            t0 = time.perf_counter()
            try:
                op._store(torch.float32)
                torch.cuda.synchronize(dev)
                built = True
            except Exception:            # e.g. out of memory on one rank: all ranks skip together
                built = False
            t_store = time.perf_counter() - t0
            if agree(built):
                for _ in range(2):
                    step_device()
                barrier()
                sev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                sev[0].record()
                for _ in range(args.steps):
                    step_device()
                sev[1].record()
                barrier()
                st = torch.tensor([sev[0].elapsed_time(sev[1]) / args.steps, t_store * 1e3, float(op.stored_bytes)],
                                  dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(st, op=dist.ReduceOp.MAX)
                stored = {"value": 2 * tot_seg / (st[0].item() / 1e3), "unit": "ray-segments/s",
                          "ms_per_step": st[0].item(), "store_ms": st[1].item(),
                          "stored_gb_per_gpu": st[2].item() / 1e9,
                          "note": "same step through the stored operator; not the headline value"}

    if rank == 0:
        ms = elapsed_ms / args.steps
        e2e_step = e2e_ms / args.steps
        value = 2 * tot_seg / (ms / 1e3)
        nvox = math.prod(full_shape[-3:]) * (full_shape[0] if dynamic else 1)
        peak, how = measured_peak()
        seg_r, rays_r = float(n_seg), float(n_rays)           # rank 0's launch
        dom = "adjoint" if t_adj >= t_fwd else "forward"
        coop = op.walker() == "cooperative"
        kname = "walk_coop_kernel" if coop else "walk_tracked_kernel"
        # DRAM bytes per ray of that kernel from the committed ncu --set full capture
        # (profiles/: a 524 288-ray launch of this workload), scaled to this launch
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
                tr = json.load(f)[kname][dom]
            traffic = tr["dram_bytes"] / tr["rays"] * rays_r if args.workload == "c5" else None
        except Exception:
            pass
        per = ADJ_BYTES if dom == "adjoint" else FWD_BYTES
        alg_bytes = per[0] * seg_r + per[1] * rays_r + per[2] * nvox
        t_dom = max(t_adj, t_fwd)
        achieved = alg_bytes / (t_dom / 1e3) / 1e9
        h2d = x_host.numel() * 4 + y_host.numel() * 4
        d2h = out_host.numel() * 4 + grad_host.numel() * 4
        line = {
            "metric": "ray-segments/s fwd+adjoint", "value": value, "unit": "ray-segments/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64 geometry / f32 volume", "data": "synthetic",
            "config": dict(desc, rays=int(tot_rays), segments_per_pass=int(tot_seg),
                           sharding=f"observations over {world} rank(s)",
                           l2="per-pass inputs (ray directions, 24 B/ray) exceed L2 for c5; "
                              "no explicit flush"),
            "rays_per_s": 2 * tot_rays / (ms / 1e3),
            "kernel_ms": {"forward": t_fwd, "adjoint": t_adj},
            "roofline": {"bound": "hbm", "kernel": f"{kname}<{dom}>",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": how,
                         "traffic_source": "ncu dram__bytes_read+write of a 524288-ray launch, scaled per ray",
                         "algorithmic_bytes": f"{per[0]} B/segment + {per[1]} B/ray + {per[2]} B/voxel"},
            "e2e": {"value": 2 * tot_seg / (e2e_step / 1e3), "unit": "ray-segments/s",
                    "ms_per_step": e2e_step, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            # per pass: cooperative walker (two launches from 4 Mi rays on: ordinary + all-cones
            # instantiation) + masked per-thread pass + sorted-cone-list pass; per-thread walker:
            # dense pass + sorted-cone-list pass
            "gpu_launches": ((8 if rays_r >= (4 << 20) else 6) if coop else 4) * args.steps,
            "walker": op.walker(),
            "clocks": clocks,
            "peak_memory_gb": peak_bytes / 1e9,
            # what the reference keeps for the same rays (regs + lens, raytracer.py:17-45) and its own
            # working-set estimator (examples/memory_usage.py:23-44), whole job
            "reference_memory_gb": reference_memory_gb(grid, tot_rays),
            "operator_build_s": t_build,
        }
        # the bounds that are actually near: one L2 gather / one L2 fp32 reduction per
        # segment against the measured peaks of tools/microbench/peaks.cu (profiles/)
        try:
            with open(os.path.join(ROOT, "profiles", "r1c_secondary_peaks.json")) as f:
                pk = json.load(f)
            line["l2_rooflines"] = {
                "forward_gathers_per_s": seg_r / (t_fwd / 1e3), "gather_peak_per_s": pk["l2_random_gather_per_s"],
                "forward_frac": seg_r / (t_fwd / 1e3) / pk["l2_random_gather_per_s"],
                "adjoint_reds_per_s": seg_r / (t_adj / 1e3), "red_peak_per_s": pk["l2_random_red_f32_per_s"],
                "adjoint_frac": seg_r / (t_adj / 1e3) / pk["l2_random_red_f32_per_s"],
                "peak_source": "tools/microbench/peaks.cu on this GPU model (profiles/r1c_secondary_peaks.json)"}
        except Exception:
            pass
        if stored is not None:
            line["stored_operator"] = stored
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline(args.workload)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("SPHRT_BENCH_WORKLOAD", "c5"))
    ap.add_argument("--obs", type=int, default=int(os.environ.get("SPHRT_BENCH_OBS", "0")) or None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
