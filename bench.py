#!/usr/bin/env python3
"""Benchmark of the forward + adjoint line-integration hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c2|c1|c4] [--obs M]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference        # the unmodified reference (baseline/_ref) on the host, bounded sample

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


# ---- CPU baseline: the reference itself (baseline/_ref) on a bounded sample ----------------------
def _reference_package():
    """The unmodified reference, installed from /root/reference into baseline/_ref;
    None if that directory did not travel."""
    path = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(path, "sph_raytracer")):
        return None
    if path not in sys.path:
        sys.path.insert(0, path)
    import sph_raytracer as R
    return R


def _reference_workload(R, name):
    """(grid, geom, sample description, volume shape) built with the REFERENCE's own classes."""
    if name == "c5":
        grid = R.SphericalGrid(shape=(200, 180, 360), size_r=(3, 25))
        gen = torch.Generator().manual_seed(0)
        theta = torch.rand(1000, generator=gen, dtype=torch.float64) * 2 * math.pi
        phi = (0.2 + 0.6 * torch.rand(1000, generator=gen, dtype=torch.float64)) * math.pi
        t, p = theta[0].item(), phi[0].item()
        g = R.ConeRectGeom((512, 512), (100 * math.cos(t) * math.sin(p), 100 * math.sin(t) * math.sin(p),
                                        100 * math.cos(p)), fov=(30, 30))
        # the whole workload needs 9.4 TB of stored operator in the reference: one observation,
        # every 4th pixel; ViewGeom with explicit per-ray starts (quirk 5 workaround)
        rays = g.rays[::4, ::4].clone()
        geom = R.ViewGeom(g.ray_starts.expand(rays.shape).clone(), rays)
        return grid, geom, "1 observation, every 4th pixel (16384 rays) of the c5 workload", tuple(grid.shape)
    if name in ("c2", "c4"):
        n_obs = 50 if name == "c2" else 20
        shape = (50, 50, 50) if name == "c2" else (n_obs, 50, 50, 50)
        grid = R.SphericalGrid(shape=shape)
        geom = sum(R.ConeCircGeom((100, 50), (5 * torch.cos(t), 5 * torch.sin(t), 1), fov=(0, 45))
                   for t in torch.linspace(0, 2 * torch.pi, n_obs))
        return grid, geom, f"the whole {name} workload ({n_obs * 5000} rays)", shape
    if name == "c1":
        grid = R.SphericalGrid(shape=(50, 50, 50))
        geom = R.ConeRectGeom((256, 256), (5, 0, 0), fov=(45, 45))
        return grid, geom, "the whole c1 workload (65536 rays)", (50, 50, 50)
    raise SystemExit(f"unknown workload {name}")


def _reference_segments(op):
    """Canonical ray-segments of a reference Operator: slots with a non-zero length, consecutive
    identical voxels merged."""
    K = op.lens.shape[-1]
    lens = op.lens.reshape(-1, K)
    regs = op.regs.reshape(3, -1, K)
    total = 0
    step = max(1, (1 << 24) // K)                 # bounded temporaries
    for a in range(0, lens.shape[0], step):
        l, r = lens[a:a + step], regs[:, a:a + step]
        mask = l > 0
        v = (r[0] * (1 << 20) + r[1]) * (1 << 20) + r[2]
        idx = torch.where(mask, torch.arange(K)[None, :], -1)
        prev = torch.cummax(idx, dim=1).values
        prev = torch.cat([torch.full((len(prev), 1), -1), prev[:, :-1]], 1)
        vprev = torch.where(prev >= 0, torch.gather(v, 1, prev.clamp(min=0)), -1)
        total += int((mask & (v != vprev)).sum())
    return total


def reference_baseline(name, reps=5):
    """The reference's own CPU path on the box's host cores: build Operator(grid, geom) (float64,
    device cpu, reference defaults), then best-of-`reps` op(x) and op(x).sum().backward()
    (reference raytracer.py:662-713)."""
    R = _reference_package()
    if R is None:
        return None
    import resource
    grid, geom, sample, vshape = _reference_workload(R, name)
    threads = torch.get_num_threads()
    t0 = time.perf_counter()
    op = R.Operator(grid, geom)
    t_build = time.perf_counter() - t0
    n_rays = int(np.prod(tuple(geom.shape)))
    nseg = _reference_segments(op)
    x = torch.rand(vshape, generator=torch.Generator().manual_seed(0)).requires_grad_(True)
    t_fwd, t_bwd = [], []
    for _ in range(max(reps, 1)):
        x.grad = None
        t0 = time.perf_counter(); y = op(x); t1 = time.perf_counter()
        y.sum().backward(); t2 = time.perf_counter()
        t_fwd.append(t1 - t0); t_bwd.append(t2 - t1)
    f, b = min(t_fwd), min(t_bwd)
    stored = (op.regs.numel() * op.regs.element_size() + op.lens.numel() * op.lens.element_size()) / 1e9
    return {"value": 2 * nseg / (f + b), "unit": "ray-segments/s", "cores": threads, "kind": "reference",
            "sample": f"{sample}; unmodified reference (baseline/_ref) on {threads} torch threads of "
                      f"{os.cpu_count()} host cores: Operator build {t_build:.2f} s ({n_rays / t_build:.0f} rays/s), "
                      f"forward {f * 1e3:.1f} ms, backward {b * 1e3:.1f} ms (best of {reps}), "
                      f"stored regs+lens {stored:.2f} GB, peak RSS {resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6:.2f} GB",
            "segments": nseg, "rays": n_rays, "ms_per_step": (f + b) * 1e3, "build_s": t_build,
            "forward_ms": f * 1e3, "backward_ms": b * 1e3, "stored_gb": stored}


def port_baseline(name):
    """Second, labelled number: the C restatement of the reference (oracle/sph_oracle.c) on the same
    kind of sample -- forward on all host threads, adjoint single-threaded."""
    from oracle import oracle as O
    O.build()
    grid, geoms, _ = workload(name, n_obs=1)
    g = geoms[0]
    rays = g.rays[::4, ::4] if name == "c5" else g.rays
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
    t0 = time.perf_counter(); O.forward(og, regs, lens, dens, nthreads=cores)
    t1 = time.perf_counter(); O.adjoint(og, regs, lens, y); t2 = time.perf_counter()
    return {"value": 2 * nseg / (t2 - t0), "unit": "ray-segments/s", "cores": cores, "kind": "port",
            "sample": f"{len(xs)} rays of one observation; C port of the reference: build {t_build:.2f} s, "
                      f"forward {(t1 - t0) * 1e3:.1f} ms ({cores} threads), adjoint {(t2 - t1) * 1e3:.1f} ms (1 thread)"}


def cpu_baseline(name, reps=5):
    cb = reference_baseline(name, reps)
    if cb is None:                          # baseline/_ref missing: the oracle port stands in, labelled
        cb = port_baseline(name)
    return cb


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
    """--impl reference: the reference's own CPU implementation as its own arm (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_baseline(args.workload, reps=max(args.steps + args.warmup, 2))
    _, _, desc = workload(args.workload, args.obs)          # the GPU arm's config; the sample is in cpu_baseline
    line = {"impl": "reference", "metric": "ray-segments/s fwd+adjoint", "value": cb["value"],
            "unit": "ray-segments/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": cb.get("ms_per_step"), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64 geometry / f32 volume", "data": "synthetic",
            "config": desc, "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "ray-segments/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if cb["kind"] == "reference" and not args.no_configs:
        # the named shapes C1, C2, C4 whole, same fields
        line["configs"] = {}
        for name in ("c1", "c2", "c4"):
            c = reference_baseline(name, reps=3)
            line["configs"][name] = {k: c[k] for k in ("value", "rays", "segments", "forward_ms", "backward_ms",
                                                        "build_s", "stored_gb", "cores")}
    print(json.dumps(line))


# ---- the GPU arm ----------------------------------------------------------------------------------
class _Flush:
    """Writes a buffer larger than L2 between timed iterations of the small configurations."""

    def __init__(self, dev, mb=256):
        self.buf = torch.empty(mb << 20, dtype=torch.uint8, device=dev)

    def __call__(self):
        self.buf.fill_(1)


def _launches():
    from sph_raytracer_b200 import _cabi
    return int(_cabi.load().sphrt_launch_count())


def _shard(grid, geoms, rank, world):
    """This rank's geometry, the grid it integrates and its block of observations / detector rows."""
    import sph_raytracer_b200 as S
    from sph_raytracer_b200.parallel import shard_bounds, shard_detector_rows, _time_slices
    n_obs = len(geoms)
    if n_obs >= world:
        lo, hi = shard_bounds(n_obs, rank, world)
        mine = geoms[lo:hi]
        geom = S.ViewGeomCollection(*mine) if len(mine) > 1 or grid.dynamic else mine[0]
        local_grid = _time_slices(grid, lo, hi) if grid.dynamic and world > 1 else grid
        return geom, local_grid, (lo, hi), f"observations over {world} rank(s)"
    coll = S.ViewGeomCollection(*geoms) if n_obs > 1 else geoms[0]
    geom, (lo, hi) = shard_detector_rows(coll, rank, world)
    return geom, grid, (lo, hi), f"detector rows over {world} rank(s) ({n_obs} observation(s) < ranks)"


def measure(name, obs, steps, warmup, world, rank, dev, flush=None, e2e=True, sampler=None, store=False,
            fused=False):
    """Device-timed forward + adjoint steps of one workload on this rank's shard (max over ranks),
    plus the same step end to end through the public API with host buffers."""
    import torch.distributed as dist
    import sph_raytracer_b200 as S
    grid, geoms, desc = workload(name, obs)
    dynamic = grid.dynamic
    geom, local_grid, (lo, hi), sharding = _shard(grid, geoms, rank, world)
    torch.cuda.synchronize(dev)
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats(dev)
    mem_base = torch.cuda.memory_allocated(dev)        # buffers of the harness itself (L2 flush, ...)
    t0 = time.perf_counter()
    op = S.Operator(local_grid, geom, device=dev, store_segments=store)
    torch.cuda.synchronize(dev)
    t_build = time.perf_counter() - t0
    n_rays = op._rays.n
    _, n_seg = op.segment_counts()
    vshape = tuple(local_grid.shape)
    x_host = torch.rand(vshape, generator=torch.Generator().manual_seed(0)).pin_memory()
    y_host = torch.rand(op._rays.shape, generator=torch.Generator().manual_seed(1)).pin_memory()
    out_host = torch.empty(op._rays.shape).pin_memory()
    grad_host = torch.empty(vshape).pin_memory()
    x, y = x_host.to(dev), y_host.to(dev)
    reduce_grad = world > 1 and not dynamic

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(warmup):
        op._apply_forward(x)
        g = op._apply_adjoint(y, vshape)
        if reduce_grad:
            dist.all_reduce(g)
    barrier()
    if sampler is not None:
        sampler.start()                   # clocks are sampled over the timed regions only
    evs = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(4)) for _ in range(steps)]
    l0 = _launches()
    for k in range(steps):
        if flush is not None:
            flush()
        a, b, c, d = evs[k]
        a.record()
        out = op._apply_forward(x)
        b.record()
        g = op._apply_adjoint(y, vshape)
        c.record()
        if reduce_grad:
            dist.all_reduce(g)
        d.record()
    barrier()
    launches = _launches() - l0
    t_fwd = float(np.mean([a.elapsed_time(b) for a, b, c, d in evs]))
    t_adj = float(np.mean([b.elapsed_time(c) for a, b, c, d in evs]))
    if flush is None:        # back-to-back steps: one interval over all of them
        ms = evs[0][0].elapsed_time(evs[-1][3]) / steps
    else:
        ms = float(np.mean([a.elapsed_time(d) for a, b, c, d in evs]))

    # the fused normal-equations step (Operator.normal: A x and A^T (w * (A x - b)) in one walk), timed
    # like the two passes above -- an extension for retrieval loops, reported next to the headline
    t_fused = None
    if fused:
        def fused_step():
            _, gg = op.normal(x, b=y, scale=0.5)
            if reduce_grad:
                dist.all_reduce(gg)
        for _ in range(min(warmup, 2)):
            fused_step()
        barrier()
        fe = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(2)) for _ in range(steps)]
        for k in range(steps):
            if flush is not None:
                flush()
            fe[k][0].record(); fused_step(); fe[k][1].record()
        barrier()
        t_fused = float(np.mean([a.elapsed_time(b) for a, b in fe])) if flush is not None else \
            fe[0][0].elapsed_time(fe[-1][1]) / steps

    # end to end through the public API: host buffers in, host buffers out, autograd adjoint;
    # the upload of the adjoint input rides a copy stream under the forward pass, the download of
    # the integrals under the adjoint pass
    e2e_ms = None
    e2e_bad, e2e_gerr = 0.0, 0.0
    if e2e:
        copy = torch.cuda.Stream(dev)
        main = torch.cuda.current_stream(dev)

        # The volume is ONE input of the job, replicated over the ranks, and every rank's host holds
        # it (SPMD): each rank uploads 1/N of it and the slices are all-gathered over NVLink; likewise
        # the gradient, the job's ONE result besides the integrals, is reduce-scattered and each rank
        # downloads its 1/N into its own pinned buffer (the bytes over PCIe are those of ONE volume
        # each way, spread over the N links).  A volume that does not divide by N goes up from rank 0
        # and is broadcast, its gradient is reduced to rank 0.  Every rank uploads its own
        # measurements and downloads its own integrals.  (A dynamic volume is sharded by time-slice:
        # every rank moves its own slices.)
        replicated = reduce_grad
        sliced = replicated and x_host.numel() % world == 0
        per = x_host.numel() // world if sliced else 0
        xs_host = x_host.view(-1)[rank * per:(rank + 1) * per] if sliced else None
        gs_host = grad_host.view(-1)[rank * per:(rank + 1) * per] if sliced else None

        def step_e2e():
            if sliced:
                xd = torch.empty(vshape, device=dev)
                dist.all_gather_into_tensor(xd.view(-1), xs_host.to(dev, non_blocking=True))
            elif replicated and rank != 0:
                xd = torch.empty(vshape, device=dev)
            else:
                xd = x_host.to(dev, non_blocking=True)
            if replicated and not sliced:
                dist.broadcast(xd, 0)
            xd.requires_grad_(True)
            x_up = torch.cuda.Event(); x_up.record(main)
            with torch.cuda.stream(copy):
                copy.wait_event(x_up)                        # (the volume has the link to itself first)
                wd = y_host.to(dev, non_blocking=True)
                up = torch.cuda.Event(); up.record(copy)
            yy = op(xd)                                      # Operator.__call__ (autograd node)
            done = torch.cuda.Event(); done.record(main)
            with torch.cuda.stream(copy):
                copy.wait_event(done)
                out_host.copy_(yy.detach(), non_blocking=True)
            main.wait_event(up)
            yy.backward(wd)                                  # autograd adjoint
            gd = xd.grad
            if sliced:
                gpart = torch.empty(per, device=dev)
                dist.reduce_scatter_tensor(gpart, gd.view(-1))
                gs_host.copy_(gpart, non_blocking=True)
            else:
                if replicated:
                    dist.reduce(gd, 0)
                if not replicated or rank == 0:
                    grad_host.copy_(gd, non_blocking=True)
            copy.synchronize(); main.synchronize()

        step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step_e2e()
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
        # what came back is what the device-resident path computes (same kernels; the sums over the
        # ranks run in NCCL's order, hence a tolerance for the gradient)
        with torch.no_grad():
            ref_y = op(x)
            e2e_bad = 0.0 if torch.equal(out_host, ref_y.cpu()) else 1.0
            ref_g = op.T(y)
            if reduce_grad:
                dist.all_reduce(ref_g)
            ref_g = ref_g.reshape(-1)
            got = gs_host if sliced else grad_host.view(-1)
            want = ref_g[rank * per:(rank + 1) * per] if sliced else ref_g
            if sliced or not replicated or rank == 0:
                e2e_gerr = float((got - want.cpu()).abs().max() / ref_g.abs().max().clamp_min(1e-30))
    peak_bytes = torch.cuda.max_memory_allocated(dev) - mem_base
    # what of that peak is the operator's own: its persistent state (grid tables, detector frames or
    # ray directions, start cells, a stored list if any) and the scratch of one call (ray masks,
    # counters, the warp walker's records of one chunk); the rest are the caller's volumes,
    # integrals and gradients of the step (forward + adjoint + end-to-end copies)
    r = op._rays
    state = sum(t.numel() * t.element_size() for t in (r.starts, r.dirs, r.cells) if t is not None)
    if r._gen is not None:
        state += sum(v.numel() * v.element_size() for v in r._gen.values() if isinstance(v, torch.Tensor))
    state += (op._tables.nr + op._tables.ne + op._tables.na) * 8 * 8 + op.stored_bytes + op.plan_cache_bytes
    op_bytes = float(state + op._scratch_bytes)
    plan_cache_bytes = float(op.plan_cache_bytes)
    clocks = sampler.summary() if sampler is not None else None

    stats = torch.tensor([ms, e2e_ms or 0.0, t_fwd, t_adj, float(peak_bytes), t_build, t_fused or 0.0,
                          e2e_bad, e2e_gerr], dtype=torch.float64, device=dev)
    # bytes this rank moves per end-to-end step (whole job: summed over the ranks below)
    replicated_volume = world > 1 and not dynamic
    if replicated_volume and x_host.numel() % world == 0:
        vol_bytes = x_host.numel() // world * 4                  # this rank's slice
    else:
        vol_bytes = x_host.numel() * 4 if (not replicated_volume or rank == 0) else 0
    h2d = y_host.numel() * 4 + vol_bytes
    d2h = out_host.numel() * 4 + vol_bytes
    sums = torch.tensor([float(n_seg), float(n_rays), float(h2d), float(d2h)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    ms, e2e_ms, t_fwd, t_adj, peak_bytes, t_build, t_fused, e2e_bad, e2e_gerr = stats.tolist()
    tot_seg, tot_rays, h2d, d2h = sums.tolist()
    return {"op": op, "grid": grid, "local_grid": local_grid, "desc": desc, "sharding": sharding, "dynamic": dynamic,
            "ms": ms, "e2e_ms": e2e_ms if e2e else None, "t_fwd": t_fwd, "t_adj": t_adj, "peak_bytes": peak_bytes,
            "t_build": t_build, "t_fused": t_fused if fused else None, "op_bytes": op_bytes, "plan_cache_bytes": plan_cache_bytes,
            "tot_seg": tot_seg, "tot_rays": tot_rays, "n_seg": n_seg, "n_rays": n_rays,
            "launches": launches, "x": x, "y": y, "out": out, "vshape": vshape, "block": (lo, hi), "clocks": clocks,
            "h2d": h2d, "d2h": d2h,
            # the end-to-end results against the device-resident path (same kernels): integrals bit for
            # bit, gradient up to the order of its float32 sums (max |difference| / max |gradient|)
            "e2e_check": {"integrals_equal": e2e_bad == 0.0, "gradient_max_rel_diff": e2e_gerr,
                          "ok": e2e_bad == 0.0 and e2e_gerr < 1e-4} if e2e else None}


def config_line(m, world):
    """The per-configuration entry of the `configs` key (BASELINE configs[0..3])."""
    ms = m["ms"]
    return {"workload": m["desc"]["workload"], "rays": int(m["tot_rays"]), "segments_per_pass": int(m["tot_seg"]),
            "forward_ms": m["t_fwd"], "adjoint_ms": m["t_adj"], "ms_per_step": ms,
            "ray_segments_per_s": 2 * m["tot_seg"] / (ms / 1e3), "rays_per_s": 2 * m["tot_rays"] / (ms / 1e3),
            "e2e_ms_per_step": m["e2e_ms"], "peak_memory_gb": m["peak_bytes"] / 1e9,
            "operator_memory_gb": m["op_bytes"] / 1e9, "plan_cache_gb": m["plan_cache_bytes"] / 1e9,
            "reference_memory_gb": reference_memory_gb(m["grid"], m["tot_rays"]),
            "operator_build_s": m["t_build"], "sharding": m["sharding"], "walker": m["op"].walker(),
            "gpu_launches_per_step": m["launches"] / max(1, m.get("steps", 1)),
            "l2": "L2 flushed (256 MB write) between timed steps"}


def retrieval_loop(dev, iterations=100, store=False, fused=False):
    """BASELINE configs[2] (examples/static_retrieval.py:18-57): 100 Adam iterations of gd() with
    [SquareLoss, NegRegularizer] through the public API; the REFERENCE's own unchanged gd / loss /
    model classes (baseline/_ref) drive the B200 operator when they are there, else the mirrors."""
    import sph_raytracer_b200 as S
    grid, geoms, _ = workload("c2")
    op = S.Operator(grid, S.ViewGeomCollection(*geoms), device=dev, store_segments=store)
    x = torch.zeros(tuple(grid.shape), device=dev)
    x[:, 25:, :25] = 1; x[:, :25, 25:] = 1
    with torch.no_grad():
        y = op(x)
    R = _reference_package()
    if R is not None:
        from sph_raytracer.retrieval import gd
        from sph_raytracer.loss import SquareLoss, NegRegularizer
        from sph_raytracer.model import FullyDenseModel
        who = "reference gd/loss/model (baseline/_ref, unchanged)"
    else:
        from sph_raytracer_b200.retrieval import gd
        from sph_raytracer_b200.loss import SquareLoss, NegRegularizer
        from sph_raytracer_b200.model import FullyDenseModel
        who = "mirrored gd/loss/model"
    if fused:       # this package's drop-in loss class, inside whichever gd() drives the loop
        from sph_raytracer_b200.loss import FusedSquareLoss as SquareLoss
        who += " + FusedSquareLoss (one walk per iteration)"
    model = FullyDenseModel(grid)
    gd(op, y, model, lr=1e-1, loss_fns=[SquareLoss(), NegRegularizer()], num_iterations=3, progress_bar=False)
    torch.cuda.synchronize(dev)
    l0 = _launches()
    t0 = time.perf_counter()
    coeffs, _, losses = gd(op, y, model, lr=1e-1, loss_fns=[SquareLoss(), NegRegularizer()],
                           num_iterations=iterations, progress_bar=False)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    first = [float(v[0]) for v in losses.values()]
    last = [float(v[-1]) for v in losses.values()]
    return {"workload": "static_retrieval: gd() of configs[2], 100 Adam iterations, SquareLoss + NegRegularizer",
            "driver": who, "iterations": iterations, "ms_per_iteration": dt * 1e3 / iterations,
            "loss_first": first, "loss_last": last,
            "gpu_launches_per_iteration": (_launches() - l0) / iterations}


def run_gpu(args):
    import torch.distributed as dist
    import sph_raytracer_b200 as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    m = measure(args.workload, args.obs, args.steps, args.warmup, world, rank, dev, sampler=ClockSampler(local),
                fused=True)
    clocks = m["clocks"]
    op, grid, dynamic = m["op"], m["grid"], m["dynamic"]

    # self-check of the launch configuration just timed: the first observation of this rank's shard
    # launched ALONE (a different, smaller launch) must give bit-identical integrals, and the segment
    # total must equal the sum of the per-ray counts
    self_check = None
    if rank == 0 and args.workload == "c5":
        g0 = workload(args.workload, args.obs)[1][m["block"][0]]
        alone = S.Operator(m["local_grid"], g0, device=dev)
        ya = alone._apply_forward(m["x"]).reshape(-1)
        yb = m["out"].reshape(-1)[:ya.numel()]
        cnt, tot = op.segment_counts()
        self_check = {"what": "observation 0 of the timed launch == the same observation launched alone (bitwise); "
                              "segment total == sum of per-ray counts",
                      "integrals_equal": bool(torch.equal(ya, yb)),
                      "max_abs_diff": float((ya - yb).abs().max()),
                      "segments_consistent": int(cnt.sum().item()) == int(tot) == int(m["n_seg"])}
        del alone

    # Extra, not the headline: the same step with the HYBRID operator (Operator(store_segments="auto"):
    # the reference's own design -- segments traced once and kept -- for as many observations as fit
    # in this GPU's free memory at 4 bytes per segment, the other observations regenerated as in the
    # headline).  At config 5 one B200 holds about 70 % of the observations, two or more hold all.
    stored = None
    if os.environ.get("SPHRT_BENCH_STORED", "1") != "0" and not dynamic:
        def agree(flag):                 # every rank takes the same branch: minimum over the ranks
            t = torch.tensor([1.0 if flag else 0.0], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
            return t.item() > 0
        m["out"] = None
        torch.cuda.empty_cache()
        t0 = time.perf_counter()
        try:
            op._pack("auto")
            torch.cuda.synchronize(dev)
            built = True
        except Exception as exc:         # e.g. out of memory on one rank: all ranks skip together
            print(f"[bench] hybrid operator not built: {exc}", file=sys.stderr)
            built = False
        t_store = time.perf_counter() - t0
        if agree(built):
            def step():
                op._apply_forward(m["x"])
                g = op._apply_adjoint(m["y"], m["vshape"])
                if world > 1:
                    dist.all_reduce(g)
            for _ in range(2):
                step()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)
            sev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            sev[0].record()
            for _ in range(args.steps):
                step()
            sev[1].record()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)
            st = torch.tensor([sev[0].elapsed_time(sev[1]) / args.steps, t_store * 1e3, float(op.stored_bytes)],
                              dtype=torch.float64, device=dev)
            frac = torch.tensor([float(op.n_stored_rays), float(m["n_rays"])], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(st, op=dist.ReduceOp.MAX)
                dist.all_reduce(frac, op=dist.ReduceOp.SUM)
            stored = {"value": 2 * m["tot_seg"] / (st[0].item() / 1e3), "unit": "ray-segments/s",
                      "ms_per_step": st[0].item(), "store_ms": st[1].item(),
                      "stored_gb_per_gpu": st[2].item() / 1e9,
                      "stored_fraction_of_rays": frac[0].item() / max(frac[1].item(), 1.0),
                      "bytes_per_segment": 4,
                      "speedup_vs_regenerated": m["ms"] / st[0].item(),
                      "note": "same step through Operator(store_segments='auto'): packed list for the observations "
                              "that fit in free memory, the others regenerated; not the headline value"}
        op._packed = None
        torch.cuda.empty_cache()

    # the named shapes of BASELINE configs[0..3] (C1, C2, C4; C3 = the gd() loop), sharded like the
    # headline; small, so every timed step starts from a flushed L2
    configs = None
    if not args.no_configs and args.workload == "c5":
        configs = {}
        flush = _Flush(dev)
        for name in ("c1", "c2", "c4"):
            steps = 20
            cm = measure(name, None, steps, 5, world, rank, dev, flush=flush, fused=True)
            cm["steps"] = steps
            if rank == 0:
                configs[name] = config_line(cm, world)
                configs[name]["fused_normal_step_ms"] = cm["t_fused"]
            del cm
            # the same configuration through Operator(store_segments="auto") (all of it fits: the
            # segments are traced once at construction and streamed from a 4 B/segment list)
            hm = measure(name, None, steps, 5, world, rank, dev, flush=flush, e2e=False, store="auto")
            if rank == 0:
                configs[name]["stored_operator"] = {
                    "forward_ms": hm["t_fwd"], "adjoint_ms": hm["t_adj"], "ms_per_step": hm["ms"],
                    "ray_segments_per_s": 2 * hm["tot_seg"] / (hm["ms"] / 1e3),
                    "operator_build_s": hm["t_build"], "stored_gb": hm["op"].stored_bytes / 1e9,
                    "stored_fraction_of_rays": hm["op"].n_stored_rays / max(1, hm["op"]._rays.n),
                    "peak_memory_gb": hm["peak_bytes"] / 1e9}
            del hm
        if world == 1:
            configs["c3"] = retrieval_loop(dev)
            st = retrieval_loop(dev, store="auto")
            configs["c3"]["stored_operator"] = {k: st[k] for k in ("ms_per_iteration", "loss_last",
                                                                   "gpu_launches_per_iteration")}
            fu = retrieval_loop(dev, fused=True)
            configs["c3"]["fused_square_loss"] = {k: fu[k] for k in ("driver", "ms_per_iteration", "loss_last",
                                                                     "gpu_launches_per_iteration")}

    if rank == 0:
        ms, e2e_step = m["ms"], m["e2e_ms"]
        tot_seg, tot_rays = m["tot_seg"], m["tot_rays"]
        value = 2 * tot_seg / (ms / 1e3)
        full_shape = tuple(grid.shape)
        nvox = math.prod(full_shape[-3:]) * (full_shape[0] if dynamic else 1)
        peak, how = measured_peak()
        seg_r, rays_r = float(m["n_seg"]), float(m["n_rays"])           # rank 0's launch
        t_fwd, t_adj = m["t_fwd"], m["t_adj"]
        dom = "adjoint" if t_adj >= t_fwd else "forward"
        coop = op.walker() == "cooperative"
        kname = "walk_warp_kernel" if coop else "walk_tracked_kernel"
        prof = {}
        try:
            with open(os.path.join(ROOT, "profiles", "r2_kernel_profile.json")) as f:
                prof = json.load(f)
        except Exception:
            pass
        # DRAM bytes of that kernel from the committed ncu --set full capture: the per-ray part scales
        # with this launch, the one-off read of the L2-resident volume does not
        traffic = None
        tr = prof.get("traffic", {}).get(dom)
        if tr and args.workload == "c5":
            traffic = (tr["dram_bytes"] - tr["volume_bytes"]) / tr["rays"] * rays_r + tr["volume_bytes"]
        per = ADJ_BYTES if dom == "adjoint" else FWD_BYTES
        alg_bytes = per[0] * seg_r + per[1] * rays_r + per[2] * nvox
        t_dom = max(t_adj, t_fwd)
        achieved = alg_bytes / (t_dom / 1e3) / 1e9
        line = {
            "metric": "ray-segments/s fwd+adjoint", "value": value, "unit": "ray-segments/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64 geometry / f32 volume", "data": "synthetic",
            "config": dict(m["desc"], rays=int(tot_rays), segments_per_pass=int(tot_seg), sharding=m["sharding"],
                           l2="per-pass inputs (ray directions, 24 B/ray) exceed L2 for c5; no explicit flush"),
            "rays_per_s": 2 * tot_rays / (ms / 1e3),
            "kernel_ms": {"forward": t_fwd, "adjoint": t_adj},
            "roofline": {"bound": "hbm", "kernel": f"{kname}<{dom}>",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": how,
                         "traffic_source": "ncu dram__bytes_read+write of one launch (profiles/r2_kernel_profile.json): "
                                           "per-ray part scaled to this launch + one read of the volume",
                         "algorithmic_bytes": f"{per[0]} B/segment + {per[1]} B/ray + {per[2]} B/voxel"},
            "e2e": {"value": 2 * tot_seg / (e2e_step / 1e3), "unit": "ray-segments/s",
                    "ms_per_step": e2e_step, "h2d_bytes_per_step": int(m["h2d"]), "d2h_bytes_per_step": int(m["d2h"]),
                    "check": m["e2e_check"],
                    "path": "Operator.__call__ + autograd backward, pinned host buffers, copies overlapped on a copy stream; "
                            "bytes are the whole job's: with N > 1 every rank uploads 1/N of the replicated volume "
                            "(all-gathered over NVLink) and downloads 1/N of the reduce-scattered gradient, and "
                            "moves its own measurements and integrals; results checked against the "
                            "device-resident path after the timed steps"},
            # counted by the library itself (sphrt_launch_count): every kernel it launched in the timed steps
            "gpu_launches": int(m["launches"]),
            "walker": op.walker(),
            # how that kernel was chosen: timed against the per-thread walker once per problem shape at
            # construction (Operator._autotune; part of operator_build_s), reused afterwards
            "autotune": getattr(op, "autotune", None),
            "clocks": clocks,
            "peak_memory_gb": m["peak_bytes"] / 1e9,
            # the operator's own part of it (state + scratch of one call; rank 0), the rest being the
            # step's volumes, integrals and gradients incl. the end-to-end copies
            "operator_memory_gb": m["op_bytes"] / 1e9,
            # ... of which the optional plan cache (80 B per ray, kept after the first call so that later
            # calls skip the planning kernel; SPHRT_PLAN_CACHE=0 or less than ~7x its size of free
            # memory: not allocated, ~5 % slower)
            "plan_cache_gb": m["plan_cache_bytes"] / 1e9,
            # what the reference keeps for the same rays (regs + lens, raytracer.py:17-45) and its own
            # working-set estimator (examples/memory_usage.py:23-44), whole job
            "reference_memory_gb": reference_memory_gb(grid, tot_rays),
            "operator_build_s": m["t_build"],
        }
        if self_check is not None:
            line["self_check"] = self_check
        # the bound that actually binds: instruction issue.  thread-instructions per segment of the
        # dominant kernel (ncu, profiles/) x segments / (SMs x 4 schedulers x 32 lanes x clock)
        ti = prof.get("thread_inst_per_segment", {}).get(dom)
        if ti and clocks.get("sm_mhz"):
            issue_peak = 148 * 4 * 32 * clocks["sm_mhz"] * 1e6
            line["issue_roofline"] = {"thread_inst_per_segment": ti, "thread_inst_per_s": ti * seg_r / (t_dom / 1e3),
                                      "peak_thread_inst_per_s": issue_peak,
                                      "frac": ti * seg_r / (t_dom / 1e3) / issue_peak,
                                      "source": "ncu smsp__thread_inst_executed of the dominant kernel (profiles/r2_kernel_profile.json)"}
        # ... and the unit that is busier still: the LSU data pipe (shared-memory wavefronts of the list
        # merges plus the gathers / reductions), as ncu measured it on the profiled launch
        lsu = prof.get("lsu_data_pipe_pct")
        if lsu:
            line["lsu_data_pipe"] = {"busy_frac": {k: lsu[k] / 100 for k in ("forward", "adjoint")},
                                     "shared_memory_share": {k: lsu["of_which_shared_memory"][k] / 100
                                                             for k in ("forward", "adjoint")},
                                     "source": "ncu l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed, "
                                               "16 config-5 observations (profiles/r2_kernel_profile.json); "
                                               "not re-measured by this run"}
        # the bandwidth-side bounds that are near: one L2 gather / one L2 fp32 reduction per
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
        if m["t_fused"]:
            line["fused_normal_step"] = {
                "ms_per_step": m["t_fused"], "speedup_vs_two_passes": ms / m["t_fused"],
                "equivalent_ray_segments_per_s": 2 * tot_seg / (m["t_fused"] / 1e3),
                "note": "Operator.normal(x, b): A x and A^T (w * (A x - b)) in ONE walk of the rays (what a "
                        "SquareLoss + autograd iteration needs), gradient all-reduced like the adjoint's; an "
                        "extension for retrieval loops, not the headline value"}
        if stored is not None:
            line["stored_operator"] = stored
        if configs is not None:
            line["configs"] = configs
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
    ap.add_argument("--no-configs", action="store_true", help="skip the configs[0..3] entries (C1-C4)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
