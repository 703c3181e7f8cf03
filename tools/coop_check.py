#!/usr/bin/env python3
"""Quick GPU check of the warp-cooperative walker against the golden fixtures and against
the per-thread walker (debug aid; the formal tests live in tests/test_gpu_coop.py)."""
import os, sys, types, time
import numpy as np, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sph_raytracer_b200 as S
from helpers import OPERATOR_CASES, load_case, case_density, ragged_equal, canonical_from_events
import bench


def geom(xs, dirs, shape):
    xs, dirs = torch.as_tensor(xs).reshape(*shape, 3), torch.as_tensor(dirs).reshape(*shape, 3)
    return types.SimpleNamespace(ray_starts=xs, rays=dirs, shape=tuple(shape))


def golden():
    for name in OPERATOR_CASES:
        c = load_case(name)
        dens = case_density(c)
        dyn = name == "c4_dynamic"
        g = types.SimpleNamespace(r_b=c.grid.r_b, e_b=c.grid.e_b, a_b=c.grid.a_b, dynamic=dyn)
        g.shape = ((dens.shape[0],) if dyn else ()) + tuple(c.nvox)
        shape = tuple(int(v) for v in c.ray_shape)
        op = S.Operator(g, geom(c.xs, c.dirs, shape))
        vox, lens, count, stats = op.events()
        torch.cuda.synchronize()
        max_seg = max(int(c.count.max()), 1) + 8
        cc, cv, cl = canonical_from_events(vox.cpu().numpy(), lens.cpu().numpy(), count.cpu().numpy(), max_seg)
        handled = cc >= 0
        degenerate = name.startswith("chords_") or name == "hemisphere"
        mask = c.stable & handled & (not degenerate)
        n, bad = ragged_equal(c, cc, cv, cl, mask=mask)
        msg = f"{name:22s} rays {len(cc):6d} stats(coop,listed,handed,failed)={stats} checked {n} bad {len(bad)}"
        # forward / adjoint through the forced cooperative path vs the per-thread path
        if dens is not None:
            x = torch.from_numpy(dens).cuda()
            os.environ["SPHRT_WALKER"] = "tracked"; y0 = op._apply_forward(x)
            os.environ["SPHRT_WALKER"] = "coop"; y1 = op._apply_forward(x)
            w = torch.rand(y0.shape, device="cuda", dtype=y0.dtype)
            os.environ["SPHRT_WALKER"] = "tracked"; g0 = op._apply_adjoint(w, x.shape)
            os.environ["SPHRT_WALKER"] = "coop"; g1 = op._apply_adjoint(w, x.shape)
            os.environ.pop("SPHRT_WALKER")
            st = torch.from_numpy(np.broadcast_to(c.stable.reshape(shape), y0.shape[-len(shape):]).copy()).cuda()
            dy = ((y0 - y1).abs() * st).max().item()
            msg += f"  |dy| {dy:.2e} (max y {y0.abs().max().item():.2f})  |dg| {(g0 - g1).abs().max().item():.2e} (max g {g0.abs().max().item():.2f})"
        print(msg, flush=True)
        if len(bad):
            i = bad[0]
            print("   first bad ray", i, "golden count", c.count[i], "got", cc[i])
            k = c.count[i]
            gv = c.vox[c.offsets[i]:c.offsets[i + 1]]
            for j in range(min(k, max_seg)):
                if not np.array_equal(gv[j], cv[i, j]):
                    print("   first diff at seg", j, "golden", gv[j], c.seg_len[c.offsets[i] + j], "got", cv[i, j], cl[i, j]); break


def c5(n_obs=1, stride=4):
    grid, geoms, _ = bench.workload("c5", n_obs)
    g = geoms[0]
    rays = g.rays[::stride, ::stride].contiguous()
    gm = types.SimpleNamespace(ray_starts=g.ray_starts, rays=rays, shape=tuple(rays.shape[:-1]))
    op = S.Operator(grid, gm)
    vox, lens, count, stats = op.events()
    sv, sl, sc = op.segments()
    n = rays.shape[0] * rays.shape[1]
    max_seg = sv.shape[-2]
    cc, cv, cl = canonical_from_events(vox.cpu().numpy(), lens.cpu().numpy(), count.cpu().numpy(), max_seg)
    sv, sl, sc = sv.reshape(n, max_seg, 3).cpu().numpy(), sl.reshape(n, max_seg).cpu().numpy(), sc.reshape(n).cpu().numpy()
    bad = 0; first = None
    for i in range(n):
        if cc[i] < 0: continue
        k = sc[i]
        ok = cc[i] == k and np.array_equal(cv[i, :k], sv[i, :k]) and np.allclose(cl[i, :k], sl[i, :k], rtol=1e-9, atol=1e-12)
        if not ok:
            bad += 1
            if first is None: first = i
    print(f"c5 tile stride {stride}: rays {n} stats {stats} handled {(cc >= 0).sum()} mismatching vs per-thread walker {bad} first {first}")
    if first is not None:
        i = first; k = sc[i]; print("  counts", cc[i], k)
        for j in range(k):
            if not np.array_equal(cv[i, j], sv[i, j]) or abs(cl[i, j] - sl[i, j]) > 1e-9:
                print("  seg", j, "coop", cv[i, j], cl[i, j], "tracked", sv[i, j], sl[i, j]); break


def timing(spec="c5:2"):
    for mode in ("tracked", "coop"):
        os.environ["SPHRT_WALKER"] = mode
        os.environ["SPHRT_LIB"] = mode       # label only
        import importlib
        sys.argv = ["kbench", spec]
        os.environ.pop("SPHRT_LIB")
        import tools.kbench as kb
        kb.run(spec)
    os.environ.pop("SPHRT_WALKER")


if __name__ == "__main__":
    what = sys.argv[1:] or ["golden", "c5", "time"]
    if "golden" in what: golden()
    if "c5" in what: c5()
    if "time" in what:  # kernel timings, both walkers
        sys.path.insert(0, ROOT)
        timing("c5:2"); timing("c2")
