"""Line-integration operator on a spherical grid -- B200 implementation.

Mirror of the reference's ``sph_raytracer/raytracer.py`` public surface: the class
``Operator`` (constructor signature, ``__call__``, ``T``, ``grid/geom/device/dynamic``
attributes, ``repr``; reference ``raytracer.py:647-755``) and ``find_starts``
(``:605-644``).  Where the reference precomputes and stores, per ray, every candidate
crossing as ``regs (3,*rays,K) int64`` / ``lens (*rays,K) float64`` (``trace_indices``,
``:48-230``) and then gathers (``:710-713``), this operator keeps only the rays on the
GPU and regenerates the crossings inside fused CUDA kernels on every call
(``csrc/sphrt_kernels.cuh`` through the C ABI of ``include/sphrt.h``).

There is no CPU implementation here.  Host tensors are accepted (they are copied to the
GPU and the result is copied back, so reference-style CPU callers keep working), but
the arithmetic always runs in the CUDA extension; if it is missing, construction fails.
"""
import math
import types

import torch

from . import _cabi
from .geometry import ViewGeomCollection, ConeRectGeom, ConeCircGeom

__all__ = ["Operator", "find_starts", "find_starts_cuda", "r_torch", "e_torch", "a_torch", "DEVICE", "FTYPE", "ITYPE"]

# module defaults, same names as the reference (raytracer.py:12-15)
DEVICE = "cuda"
PDEVICE = "cpu"
FTYPE = torch.float64
ITYPE = torch.int64

_TAU = float.fromhex("0x1.4f8b588e368f4p-17")      # isclose threshold, raytracer.py:233-246
_DTYPES = {torch.float32: _cabi.F32, torch.float64: _cabi.F64}


def _compute_device(device):
    """CUDA device that runs the kernels for a user-facing ``device`` argument."""
    dev = torch.device(device)
    if dev.type == "cuda":
        return torch.device("cuda", torch.cuda.current_device() if dev.index is None else dev.index)
    if not torch.cuda.is_available():
        raise RuntimeError(
            "sph_raytracer_b200 needs a CUDA device: the operator has no CPU implementation")
    return torch.device("cuda", torch.cuda.current_device())


class _GridTables:
    """Boundary vectors and host-computed trigonometric tables of a SphericalGrid on the GPU.

    cos/sin are evaluated with torch on the host exactly like the reference does
    (raytracer.py:373-375, 452, 505-506) so that both sides see identical table values.
    """

    def __init__(self, grid, device):
        r_b = torch.asarray(grid.r_b, dtype=torch.float64).cpu().reshape(-1)
        e_b = torch.asarray(grid.e_b, dtype=torch.float64).cpu().reshape(-1)
        a_b = torch.asarray(grid.a_b, dtype=torch.float64).cpu().reshape(-1)
        self.r_b, self.e_b, self.a_b = r_b, e_b, a_b
        self.nr, self.ne, self.na = len(r_b) - 1, len(e_b) - 1, len(a_b) - 1
        if min(self.nr, self.ne, self.na) < 1:
            raise ValueError("grid needs at least one voxel along every axis")
        cos_e = torch.cos(e_b)
        c2 = (cos_e * cos_e)[(math.pi / 2 - e_b).abs() >= _TAU]
        c2_sorted = torch.sort(c2).values
        parts = [r_b, e_b, cos_e, a_b, torch.cos(a_b), torch.sin(a_b), c2_sorted]
        self.packed = torch.cat(parts).to(device)
        base, ptrs = self.packed.data_ptr(), []
        off = 0
        for p in parts:
            ptrs.append(base + 8 * off)
            off += len(p)
        self.c = _cabi.Grid(self.nr, self.ne, self.na, len(c2_sorted), *ptrs)
        self.nvox = self.nr * self.ne * self.na


def _compact(t, full_shape):
    """Represent ``t`` (broadcastable to ``full_shape + (3,)``) as (m, 3) rows plus the number
    of consecutive rays sharing a row, without materialising the broadcast."""
    lead = len(full_shape) - (t.dim() - 1)
    shape = (1,) * lead + tuple(t.shape[:-1])
    k = len(full_shape)
    while k > 0 and shape[k - 1] == 1:
        k -= 1
    if all(shape[i] == full_shape[i] for i in range(k)):
        per = math.prod(full_shape[k:])
        return t.reshape(-1, 3).contiguous(), max(per, 1)
    return t.broadcast_to(tuple(full_shape) + (3,)).reshape(-1, 3).contiguous(), 1


def _start_cells(r_b, e_b, a_b, starts):
    """The reference's binning of start points (cart2sph + find_starts, raytracer.py:555-578,
    605-644) evaluated on the host with torch -- the same library the reference uses, so that a
    start sitting on a cone or half-plane to the last bit lands in the bin the reference puts it
    in (atan2 differs in the last ulp between libms).  Returns int32
    (m, 8): geometric bins r,e,a, reference states r,e,a, on-boundary flag, 0."""
    x, y, z = starts.unbind(-1)
    xy_sq = x ** 2 + y ** 2
    sph = (torch.sqrt(xy_sq + z ** 2), torch.arctan2(torch.sqrt(xy_sq), z), torch.arctan2(y, x))
    cells = torch.zeros((starts.shape[0], 8), dtype=torch.int32)
    near = torch.zeros(starts.shape[0], dtype=torch.bool)
    for k, (b, v) in enumerate(zip((r_b, e_b, a_b), sph)):
        n = len(b) - 1
        v = v.contiguous()
        bins = torch.searchsorted(b, v, right=True) - 1
        bins = torch.where(v == b[-1], n - 1, bins)
        cells[:, k] = bins
        cells[:, 3 + k] = torch.where(bins == n, -1, bins)
        if k > 0:      # within 1e-9 of a boundary of its elevation / azimuth bin
            lo, hi = bins.clamp(0, n), (bins + 1).clamp(0, n)
            near |= ((v - b[lo]).abs() < 1e-9) | ((v - b[hi]).abs() < 1e-9)
    cells[:, 6] = near.to(torch.int32)
    return cells


def _pinhole_generator(geom):
    """Analytic description of a pinhole detector or a collection of them: per
    observation the frame (pos, lookdir, u, v) the reference's ``rays`` property starts from, and
    the float32 pixel-coordinate tables it multiplies them with (reference geometry.py:494-508
    rectangular, :557-568, 573-582 polar).  The kernels then form every direction themselves,
    bit-identical to ``geom.rays``, and ``geom.rays`` is never materialised.  None if ``geom`` is
    not made of such detectors (or they differ in shape)."""
    obs = list(geom.geoms) if isinstance(geom, ViewGeomCollection) else [geom]
    if not obs or not all(type(g) in (ConeRectGeom, ConeCircGeom) for g in obs):
        return None
    if len({(type(g), tuple(g.shape)) for g in obs}) != 1 or len(obs[0].shape) != 2:
        return None
    polar = type(obs[0]) is ConeCircGeom
    H, W = (int(v) for v in obs[0].shape)
    frames, ta, tb, tc = [], [], [], []
    for g in obs:
        u = torch.cross(g.lookdir, g.updir, dim=-1)
        frames.append(torch.cat([g.pos.reshape(3), g.lookdir.reshape(3), u.reshape(3), g.updir.reshape(3)]))
        if polar:
            ta.append(g.r.to(torch.float32)); tb.append(torch.cos(g.theta)); tc.append(torch.sin(g.theta))
        else:
            ta.append(torch.linspace(-g._half_extent(0), g._half_extent(0), H))
            tb.append(torch.linspace(-g._half_extent(1), g._half_extent(1), W))
    if any(t.dtype != torch.float32 for t in ta + tb + tc):
        return None
    same = all(torch.equal(t, ta[0]) for t in ta) and all(torch.equal(t, tb[0]) for t in tb) and \
        all(torch.equal(t, tc[0]) for t in tc)
    keep = (lambda ts: ts[0][None]) if same else torch.stack
    lead = (len(obs),) if isinstance(geom, ViewGeomCollection) else ()
    return {"kind": 2 if polar else 1, "H": H, "W": W, "per_obs": 0 if same else 1,
            "frames": torch.stack(frames).to(torch.float64).contiguous(),
            "ta": keep(ta).contiguous(), "tb": keep(tb).contiguous(),
            "tc": keep(tc).contiguous() if polar else None,
            "shape": lead + (H, W)}


class _Rays:
    """Ray origins and directions of a view geometry on the GPU (float64).

    Pinhole detectors (ConeRectGeom / ConeCircGeom and collections of them) are described
    analytically: 96 bytes per observation plus the pixel-coordinate tables instead of 24 bytes per
    ray; SPHRT_RAYGEN=0 uploads explicit directions as for any other geometry."""

    def __init__(self, geom, device, tables=None):
        import os
        gen = _pinhole_generator(geom) if os.environ.get("SPHRT_RAYGEN", "1") != "0" else None
        self.cells = None
        if gen is not None:
            self.shape = tuple(gen["shape"])
            self.n = math.prod(self.shape)
            s = gen["frames"][:, :3].contiguous()
            self.rays_per_start, self.rays_per_dir = gen["H"] * gen["W"], 1
            self.starts, self.dirs = s.to(device), None
            self._gen = {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in gen.items()}
        else:
            starts = torch.asarray(geom.ray_starts, dtype=torch.float64)
            dirs = torch.asarray(geom.rays, dtype=torch.float64)
            # the reference broadcasts the smaller of the two to the larger (raytracer.py:77-80)
            full = tuple(starts.shape if starts.numel() > dirs.numel() else dirs.shape)[:-1]
            if max(starts.dim(), dirs.dim()) - 1 > len(full):
                raise ValueError(f"ray_starts {tuple(starts.shape)} and rays {tuple(dirs.shape)} "
                                 "cannot be broadcast to a common detector shape")
            self.shape = tuple(full)
            self.n = math.prod(self.shape)
            s, self.rays_per_start = _compact(starts, self.shape)
            d, self.rays_per_dir = _compact(dirs, self.shape)
            self.starts, self.dirs = s.to(device), d.to(device)
            self._gen = None
        if tables is not None:
            self.cells = _start_cells(tables.r_b, tables.e_b, tables.a_b, s.cpu()).to(device)
        self.c = self.struct(self.n)

    def struct(self, n_rays, walker=0):
        """``struct sphrt_rays`` of the first ``n_rays`` rays."""
        g = self._gen
        ptr = lambda t: t.data_ptr() if t is not None else None
        return _cabi.Rays(n_rays, self.starts.data_ptr(), self.rays_per_start,
                          ptr(self.dirs), self.rays_per_dir, ptr(self.cells), walker,
                          g["kind"] if g else 0, g["H"] if g else 0, g["W"] if g else 0, g["per_obs"] if g else 0,
                          ptr(g["frames"]) if g else None, ptr(g["ta"]) if g else None,
                          ptr(g["tb"]) if g else None, ptr(g["tc"]) if g else None)

    @property
    def unit(self):
        """Rays per observation: the granularity at which the ray list can be split (a prefix /
        suffix of whole observations is again a valid ``struct sphrt_rays``)."""
        if self._gen is not None:
            return self._gen["H"] * self._gen["W"]
        u = max(self.rays_per_start, self.rays_per_dir)
        return u if u % self.rays_per_start == 0 and u % self.rays_per_dir == 0 and self.n % u == 0 else self.n

    def struct_from(self, lo, walker=0):
        """``struct sphrt_rays`` of the rays ``lo ..`` (``lo`` a multiple of ``unit``)."""
        if lo % self.unit or not 0 <= lo <= self.n:
            raise ValueError(f"ray {lo} is not the first ray of an observation ({self.unit} rays each)")
        g = self._gen
        at = lambda t, k: (t.data_ptr() + k * t.element_size()) if t is not None else None
        s0 = lo // self.rays_per_start
        if g:
            o = lo // self.unit
            per = g["per_obs"]
            return _cabi.Rays(self.n - lo, at(self.starts, 3 * s0), self.rays_per_start, None, self.rays_per_dir,
                              at(self.cells, 8 * s0), walker, g["kind"], g["H"], g["W"], per,
                              at(g["frames"], 12 * o), at(g["ta"], per * o * g["H"]),
                              at(g["tb"], per * o * g["W"]), at(g["tc"], per * o * g["W"]))
        return _cabi.Rays(self.n - lo, at(self.starts, 3 * s0), self.rays_per_start,
                          at(self.dirs, 3 * (lo // self.rays_per_dir)), self.rays_per_dir,
                          at(self.cells, 8 * s0), walker, 0, 0, 0, 0, None, None, None, None)

    def directions(self, lib, stream):
        """Every ray's direction as the kernels see it, ``(*shape, 3)`` float64 (parity export)."""
        out = torch.empty((self.n, 3), dtype=torch.float64, device=self.starts.device)
        with torch.cuda.device(out.device):
            _cabi.check(lib.sphrt_export_rays(self.c, out.data_ptr(), stream))
        return out.reshape(self.shape + (3,))


_AUTOTUNE_CACHE = {}


class _LineIntegral(torch.autograd.Function):
    """density -> line integrals; backward is the scatter-add adjoint (the reference gets the
    same pair implicitly from ATen's index / index_put_)."""

    @staticmethod
    def forward(ctx, density, op):
        ctx.op = op
        ctx.in_shape, ctx.in_device = density.shape, density.device
        return op._apply_forward(density)

    @staticmethod
    def backward(ctx, grad_out):
        op = ctx.op
        grad = op._apply_adjoint(grad_out, ctx.in_shape)
        return grad.to(ctx.in_device), None


class Operator:
    """Raytracing operator: ``op(density)`` integrates ``density`` along every line of sight.

    Args (same as the reference, raytracer.py:662-665):
        grid (SphericalGrid): voxel boundaries
        geom (ViewGeom): lines of sight
        dynamic (bool): kept for ``repr``; None = True for a ViewGeomCollection
        ftype: only torch.float64 geometry is implemented (the reference's float32 mode
            changes which voxels are hit)
        itype, pdevice: accepted and ignored (nothing is materialised)
        device: where callers create their tensors.  The kernels always run on a CUDA
            device (the given one, or the current one if ``device`` is a CPU device)
        debug, debug_los: print the segment list of one line of sight
        invalid: not supported (it exposes un-zeroed padding slots that do not exist here)
        _compute: False skips uploading the geometry
        store_segments (extension): False (default) regenerates the crossings in
            every call and keeps only the rays on the GPU.  True / torch.float32 / torch.float64
            traces once at construction and stores the canonical segments as a compact list
            (8 bytes per segment, 12 with float64 lengths; the reference stores 32 bytes for each
            of its K padded slots per ray, raytracer.py:675-679); ``op(x)`` / ``op.T(y)`` then
            stream that list -- for retrieval loops that apply one operator hundreds of times
            and have the memory (``op.stored_bytes``).  Same results to the precision of the
            stored lengths (float32 lengths: relative 6e-8 per segment).
            "packed" stores 4 bytes per segment (3-bit step code + 29-bit length, relative
            2.4e-7; ``include/sphrt.h``).  "auto" packs as many whole observations as fit in
            the memory budget (80 % of the free device memory, or ``SPHRT_STORE_BUDGET_GB``)
            and regenerates the others in every call: the hybrid operator
            (``op.n_stored_rays`` of ``op._rays.n`` rays come from the list).
    """

    def __init__(self, grid, geom, dynamic=False, ftype=FTYPE, itype=ITYPE, device=DEVICE,
                 pdevice=PDEVICE, debug=False, debug_los=None, invalid=False, _compute=True,
                 store_segments=False):
        self.grid = grid
        self.geom = geom
        if dynamic is None:
            dynamic = isinstance(geom, ViewGeomCollection)
        self.dynamic = dynamic
        self.ftype, self.itype, self.device = ftype, itype, device
        if ftype != torch.float64:
            raise NotImplementedError("only ftype=torch.float64 geometry is implemented")
        if invalid:
            raise NotImplementedError("invalid=True (raw padded slots) has no equivalent here")
        if not _compute:
            return
        self._lib = _cabi.load()
        self._cdev = _compute_device(device)
        self._tables = _GridTables(grid, self._cdev)
        self._rays = _Rays(geom, self._cdev, self._tables)
        self._scratch_bytes = int(self._lib.sphrt_scratch_bytes(self._tables.c, self._rays.n))
        self._stored = None
        self._packed = None
        self._pc = {}                    # plan caches: "main" (all rays), "suffix" (hybrid operator)
        self._autotune()
        if isinstance(store_segments, str):
            if store_segments not in ("packed", "auto"):
                raise ValueError(f"store_segments={store_segments!r}: expected False, True, a dtype, "
                                 "'packed' or 'auto'")
            self._pack(None if store_segments == "packed" else "auto")
        elif store_segments:
            self._store(torch.float64 if store_segments is torch.float64 else torch.float32)
        if debug:
            self._print_los(debug_los)

    # ---- launch helpers ---------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self._cdev).cuda_stream

    def _scratch(self):
        return torch.empty(max(self._scratch_bytes, 8), dtype=torch.uint8, device=self._cdev)

    def _autotune(self, max_rays=1 << 20):
        """Pick the kernel for the ordinary rays by measurement: the warp-cooperative and the
        per-thread walker give the same segments, which one is faster depends on grid size and
        on how alike neighbouring rays are.  One warm-up and two timed forward passes per
        kernel over (a prefix of) this operator's own rays; SPHRT_AUTOTUNE=0 keeps the library's
        size heuristic, SPHRT_WALKER forces a kernel."""
        import os
        import time
        self.autotune = {"walker": "library heuristic", "ms": 0.0, "cached": False}
        if os.environ.get("SPHRT_AUTOTUNE", "1") == "0" or os.environ.get("SPHRT_WALKER") or self._rays.n < 4096:
            return
        r, t = self._rays, self._tables
        # Grids on which the library's size heuristic already takes the cooperative walker (a ray meets
        # a few hundred boundaries: it is 2-3x the per-thread walker there) are not measured: the
        # choice is the same in every process -- also under a profiler, whose per-launch overhead
        # would weigh on the path with more launches.
        if int(self._lib.sphrt_walker_kind(t.c, r.n)) == 1:
            r.c.walker = 1
            self.autotune = {"walker": "cooperative", "ms": 0.0, "cached": False, "decided_by": "size heuristic"}
            return
        # one decision per (grid shape, detector shape, device): operators built again for the same
        # problem (a retrieval script, the bench's repetitions) reuse it and stay deterministic
        key = (t.nr, t.ne, t.na, tuple(r.shape), r._gen is not None, str(self._cdev))
        if key in _AUTOTUNE_CACHE:
            r.c.walker = _AUTOTUNE_CACHE[key]
            self.autotune = {"walker": "cooperative" if r.c.walker == 1 else "per-thread", "ms": 0.0, "cached": True}
            return
        t_start = time.perf_counter()
        n_probe = min(r.n, max_rays)
        if r._gen is not None:                      # whole observations
            hw = r._gen["H"] * r._gen["W"]
            n_probe = max(hw, n_probe // hw * hw)
        probe = r.struct(n_probe)
        vol = torch.ones(t.nvox, dtype=torch.float32, device=self._cdev)
        out = torch.empty(probe.n_rays, dtype=torch.float32, device=self._cdev)
        scratch = self._scratch()
        best = None
        with torch.cuda.device(self._cdev):
            for walker in (1, 2):
                probe.walker = walker
                times = []
                for k in range(3):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    _cabi.check(self._lib.sphrt_forward(t.c, probe, vol.data_ptr(), _cabi.F32, 1, 0,
                                                        out.data_ptr(), scratch.data_ptr(),
                                                        scratch.numel(), self._stream()))
                    b.record()
                    b.synchronize()
                    times.append(a.elapsed_time(b))
                if best is None or min(times[1:]) < best[0]:
                    best = (min(times[1:]), walker)
        r.c.walker = best[1]
        _AUTOTUNE_CACHE[key] = best[1]
        self.autotune = {"walker": "cooperative" if best[1] == 1 else "per-thread",
                         "ms": (time.perf_counter() - t_start) * 1e3, "cached": False}

    def _store(self, len_dtype):
        """Trace once and keep the canonical segments (offsets, flat voxel offsets, lengths)."""
        count, total = self.segment_counts()
        offsets = torch.zeros(self._rays.n + 1, dtype=torch.int64, device=self._cdev)
        torch.cumsum(count.reshape(-1), 0, out=offsets[1:])
        vox = torch.empty(max(total, 1), dtype=torch.int32, device=self._cdev)
        lens = torch.empty(max(total, 1), dtype=len_dtype, device=self._cdev)
        scratch = self._scratch()
        with torch.cuda.device(self._cdev):
            _cabi.check(self._lib.sphrt_store_segments(
                self._tables.c, self._rays.c, offsets.data_ptr(), vox.data_ptr(), lens.data_ptr(),
                _DTYPES[len_dtype], scratch.data_ptr(), scratch.numel(), self._stream()))
        self._stored = (offsets, vox, lens)
        self.n_segments = total

    def _store_budget(self):
        """Bytes the packed list may take: SPHRT_STORE_BUDGET_GB, else 80 % of the memory that is
        free (in the driver or cached by torch's allocator) when the list is about to be built."""
        import os
        env = os.environ.get("SPHRT_STORE_BUDGET_GB")
        if env:
            return int(float(env) * 1e9)
        torch.cuda.empty_cache()
        free, _ = torch.cuda.mem_get_info(self._cdev)
        return int(0.8 * free)

    def _pack(self, budget=None):
        """Pack the segments of the first observations (all of them if ``budget`` is None or they
        fit) at 4 bytes per segment; the other observations stay regenerated (hybrid operator)."""
        r, t = self._rays, self._tables
        # (the plan cache of the whole ray list is of no use to a hybrid operator: its memory goes
        # to the packed list)
        self._pc = {"main": {"buf": False, "key": None, "pending": None}}
        r.c.plan_cache, r.c.plan_cache_bytes, r.c.plan_cache_state = None, 0, 0
        torch.cuda.empty_cache()             # (hand its segment back before anything is carved out of it)
        count = torch.empty(r.n, dtype=torch.int32, device=self._cdev)
        scratch = self._scratch()
        with torch.cuda.device(self._cdev):
            _cabi.check(self._lib.sphrt_pack_count(t.c, r.c, count.data_ptr(), None,
                                                   scratch.data_ptr(), scratch.numel(), self._stream()))
        offsets = torch.zeros(r.n + 1, dtype=torch.int64, device=self._cdev)
        torch.cumsum(count, 0, out=offsets[1:])
        del count
        n_s = r.n
        if budget == "auto":
            budget = self._store_budget() - 8 * (r.n + 1)            # (the offsets are cut out of a copy)
        if budget is not None:
            unit = r.unit
            ends = offsets[unit::unit]                                   # words up to each observation's end
            cost = 4 * ends + 8 * (torch.arange(1, ends.numel() + 1, device=self._cdev) * unit + 1)
            n_s = int((cost <= budget).sum().item()) * unit
        if n_s < r.n:
            offsets = offsets[:n_s + 1].clone()
        total = int(offsets[-1].item())
        words = torch.empty(max(total, 1), dtype=torch.int32, device=self._cdev)
        if n_s > 0:
            with torch.cuda.device(self._cdev):
                _cabi.check(self._lib.sphrt_pack_segments(t.c, r.struct(n_s, r.c.walker), offsets.data_ptr(),
                                                          words.data_ptr(), scratch.data_ptr(),
                                                          scratch.numel(), self._stream()))
        self._packed = (offsets, words, n_s)
        self._suffix = r.struct_from(n_s, r.c.walker) if n_s < r.n else None
        self._suffix_scratch_bytes = int(self._lib.sphrt_scratch_bytes(t.c, r.n - n_s)) if n_s < r.n else 0

    @property
    def n_stored_rays(self):
        """Rays served from a stored list (the others are regenerated in every call)."""
        if self._packed is not None:
            return self._packed[2]
        return self._rays.n if self._stored is not None else 0

    @property
    def stored_bytes(self):
        """Device memory held by the stored segment list (0 when segments are regenerated)."""
        if self._packed is not None:
            return sum(t.numel() * t.element_size() for t in self._packed[:2])
        return 0 if self._stored is None else sum(t.numel() * t.element_size() for t in self._stored)

    def _hybrid_call(self, forward, src, dst, n_vol, rays_per_vol):
        """Packed prefix + regenerated suffix of the ray list.  The C ABI takes dense
        ``(n_vol, n_rays)`` blocks, so each part works on its own block of the result (forward) or
        of the weights (adjoint; both parts accumulate into the same gradient)."""
        offsets, words, n_s = self._packed
        n, t = self._rays.n, self._tables
        packed = self._lib.sphrt_forward_packed if forward else self._lib.sphrt_adjoint_packed
        traced = self._lib.sphrt_forward if forward else self._lib.sphrt_adjoint
        rays_side = dst if forward else src                 # the tensor indexed by ray
        if rays_side.dim() == 2 and n_vol == 1:
            rays_side = rays_side.reshape(-1)
        flat = rays_side.dim() == 1
        parts = []
        with torch.cuda.device(self._cdev):
            for stored_part, lo, hi in ((True, 0, n_s), (False, n_s, n)):
                if lo == hi:
                    continue
                whole = lo == 0 and hi == n
                if whole:
                    blk = rays_side
                elif flat:
                    blk = rays_side[lo:hi]                      # contiguous slice, used in place
                else:
                    blk = torch.empty((n_vol, hi - lo), dtype=rays_side.dtype, device=self._cdev) if forward \
                        else rays_side[:, lo:hi].contiguous()
                vol = dst if not forward else src
                nv = n_vol
                if rays_per_vol > 0 and lo > 0:                 # observation o reads time-slice o
                    o0 = lo // rays_per_vol
                    vol, nv = vol.reshape(n_vol, -1)[o0:], n_vol - o0
                a, b = (vol, blk) if forward else (blk, vol)
                if stored_part:
                    _cabi.check(packed(offsets.data_ptr(), words.data_ptr(), hi, t.ne, t.na, a.data_ptr(),
                                       _DTYPES[a.dtype], nv, rays_per_vol, t.nvox, b.data_ptr(), self._stream()))
                else:
                    scratch = torch.empty(max(self._suffix_scratch_bytes, 8), dtype=torch.uint8, device=self._cdev)
                    self._plan_cache_begin("suffix")
                    _cabi.check(traced(t.c, self._suffix, a.data_ptr(), _DTYPES[a.dtype], nv, rays_per_vol,
                                       b.data_ptr(), scratch.data_ptr(), scratch.numel(), self._stream()))
                    self._plan_cache_end("suffix")
                parts.append((lo, hi, blk, whole))
        if forward and not flat:
            for lo, hi, blk, whole in parts:
                if not whole:
                    dst[:, lo:hi] = blk

    def _plan_cache_begin(self, which="main"):
        """Plan cache of the warp walker (``sphrt_rays.plan_cache``): the planner's 80-byte record per
        ray and its two bits per ray are the same in every call, so they are kept after the first
        one if that costs less than 15 % of the free device memory (SPHRT_PLAN_CACHE=0: never).
        ``which``: the operator's whole ray list, or the regenerated suffix of a hybrid operator."""
        import os
        c = self._pc.setdefault(which, {"buf": None, "key": None, "pending": None})
        rays = self._rays.c if which == "main" else self._suffix
        n = self._rays.n if which == "main" else int(rays.n_rays)
        if c["buf"] is False or os.environ.get("SPHRT_PLAN_CACHE", "1") == "0":
            rays.plan_cache_state = 0
            return
        key = tuple(os.environ.get(k) for k in ("SPHRT_WARP_CHUNK", "SPHRT_WALKER", "SPHRT_COOP_GEN", "SPHRT_AUTOTUNE"))
        if c["buf"] is None:
            nbytes = int(self._lib.sphrt_plan_cache_bytes(self._tables.c, n))
            if nbytes == 0 or self.walker() != "cooperative" or \
                    nbytes > 0.15 * torch.cuda.mem_get_info(self._cdev)[0]:
                c["buf"] = False
                return
            c["buf"] = torch.empty(nbytes, dtype=torch.uint8, device=self._cdev)
            rays.plan_cache = c["buf"].data_ptr()
            rays.plan_cache_bytes = nbytes
            c["key"] = None
        if key != c["key"]:                  # first call, or the environment picked another kernel path
            c["pending"] = key               # (becomes the cache's key once the call has succeeded)
            rays.plan_cache_state = 1
        else:
            rays.plan_cache_state = 2

    def _plan_cache_end(self, which="main", ok=True):
        c = self._pc.get(which)
        if c is not None and c["pending"] is not None:
            if ok:
                c["key"] = c["pending"]
            c["pending"] = None

    @property
    def plan_cache_bytes(self):
        return sum(c["buf"].numel() for c in self._pc.values() if isinstance(c["buf"], torch.Tensor))

    def _volume_call(self, forward, src, dst, n_vol, rays_per_vol):
        if self._packed is not None:
            return self._hybrid_call(forward, src, dst, n_vol, rays_per_vol)
        if self._stored is not None:
            offsets, vox, lens = self._stored
            fn = self._lib.sphrt_forward_stored if forward else self._lib.sphrt_adjoint_stored
            with torch.cuda.device(self._cdev):
                _cabi.check(fn(offsets.data_ptr(), vox.data_ptr(), lens.data_ptr(), _DTYPES[lens.dtype],
                               self._rays.n, src.data_ptr(), _DTYPES[src.dtype], n_vol, rays_per_vol,
                               self._tables.nvox, dst.data_ptr(), self._stream()))
            return
        scratch = self._scratch()
        fn = self._lib.sphrt_forward if forward else self._lib.sphrt_adjoint
        self._plan_cache_begin()
        with torch.cuda.device(self._cdev):
            _cabi.check(fn(self._tables.c, self._rays.c, src.data_ptr(), _DTYPES[src.dtype],
                           n_vol, rays_per_vol, dst.data_ptr(), scratch.data_ptr(),
                           scratch.numel(), self._stream()))
        self._plan_cache_end()

    # ---- batching rules of Operator.__call__ (raytracer.py:703-713) ---------------------
    def _plan(self, shape):
        """(n_vol, rays_per_vol, out_shape, channel_axis) for a density of ``shape``."""
        t = self._tables
        G = self._rays.shape
        if tuple(shape[-3:]) != (t.nr, t.ne, t.na):
            raise IndexError(f"density shape {tuple(shape)} does not end in the grid shape "
                             f"{(t.nr, t.ne, t.na)}")
        if not self.grid.dynamic:
            batch = tuple(shape[:-3])
            return math.prod(batch), 0, batch + G, None
        if len(shape) != 4:
            raise IndexError("a dynamic grid takes a 4D density (time, r, e, a)")
        T, g = shape[0], len(G)
        if g < 3:                                   # every time-slice sees the same rays
            return T, 0, (T,) + (1,) * (2 - g) + G, None
        n_obs = G[g - 3]
        if n_obs == T:                              # observation o reads time-slice o
            return T, math.prod(G[g - 2:]), G, None
        if T == 1:
            return 1, 0, G, None
        if n_obs == 1:
            return T, 0, G[:g - 3] + (T,) + G[g - 2:], g - 3
        raise IndexError(f"shape mismatch: {T} time-slices cannot be broadcast against "
                         f"{n_obs} observations")

    def _apply_forward(self, density):
        if density.dtype not in _DTYPES:
            raise TypeError(f"density dtype {density.dtype} not supported (float32 / float64)")
        n_vol, rpv, out_shape, ch_axis = self._plan(density.shape)
        d = density.detach().to(self._cdev).contiguous()
        n = self._rays.n
        out = torch.empty((n,) if rpv else (n_vol, n), dtype=d.dtype, device=self._cdev)
        if n_vol > 0 and n > 0:
            self._volume_call(True, d, out, n_vol, rpv)
        if ch_axis is not None:
            G = self._rays.shape
            out = out.reshape((n_vol,) + G[:ch_axis] + G[ch_axis + 1:]).movedim(0, ch_axis)
        return out.reshape(out_shape).to(density.device)

    def _apply_adjoint(self, y, density_shape):
        n_vol, rpv, out_shape, ch_axis = self._plan(density_shape)
        if y.dtype not in _DTYPES:
            raise TypeError(f"dtype {y.dtype} not supported (float32 / float64)")
        yd = y.detach().to(self._cdev).reshape(out_shape)
        n = self._rays.n
        if ch_axis is not None:
            yd = yd.movedim(ch_axis, 0)
        yd = yd.reshape((n,) if rpv else (n_vol, n)).contiguous()
        grad = torch.zeros((max(n_vol, 0), self._tables.nr, self._tables.ne, self._tables.na),
                           dtype=yd.dtype, device=self._cdev)
        if n_vol > 0 and n > 0:
            self._volume_call(False, yd, grad, n_vol, rpv)
        return grad.reshape(density_shape)

    # ---- public API --------------------------------------------------------------------
    def __call__(self, density):
        """Line integrals of ``density``.

        static grid : density ``(*batch, N_r, N_e, N_a)`` -> ``(*batch, *geom.shape)``
        dynamic grid: density ``(T, N_r, N_e, N_a)``; with a ViewGeomCollection of length T
                      observation *o* integrates time-slice *o*            -> ``geom.shape``
        Output dtype and device follow ``density``.  Differentiable w.r.t. ``density``.
        """
        return _LineIntegral.apply(density, self)

    def T(self, line_integrations):
        """Adjoint (back-projection) of the operator: ``grid.shape`` volume from line
        integrals of shape ``geom.shape`` (reference raytracer.py:715-748).

        The reference implements the static case only and raises ``NotImplementedError`` for
        a dynamic grid; here the dynamic case is the adjoint of the dynamic forward
        (observation *o* back-projects into time-slice *o*).  Any float32 /
        float64 input is accepted (the reference needs float64)."""
        shape = tuple(self.grid.shape)
        y = line_integrations
        want = self._plan(shape)[2]
        if tuple(y.shape) != tuple(want):
            raise IndexError(f"expected line integrals of shape {tuple(want)}, got {tuple(y.shape)}")
        return self._apply_adjoint(y, shape).to(y.device)

    def normal(self, density, b=None, mask=None, scale=1.0):
        """Fused normal-equations step (extension): ``(A x, A^T (scale * mask * (A x - b)))``.

        ``density`` has the grid's shape; ``b`` and ``mask`` are broadcastable to ``geom.shape``
        (None: b = 0, mask = 1).  Both results come from ONE walk of the rays (``sphrt_normal``:
        the weight of a ray's back-projection depends only on its own integral, and its merged
        crossing lists are still in shared memory after the forward sweep) -- what one
        SquareLoss + autograd iteration of a retrieval loop needs (reference loss.py:86-90,
        retrieval.py:99-116).  No autograd; :class:`loss.FusedSquareLoss` wraps it for ``gd``.
        A stored / hybrid operator, or a grid the warp-cooperative walker does not serve, runs the
        two passes one after the other instead."""
        with torch.no_grad():
            shape = tuple(self.grid.shape)
            if tuple(density.shape) != shape:
                raise IndexError(f"density shape {tuple(density.shape)} is not the grid shape {shape}")
            n_vol, rpv, out_shape, ch_axis = self._plan(shape)
            d = density.detach().to(self._cdev).contiguous()
            if d.dtype not in _DTYPES:
                raise TypeError(f"density dtype {d.dtype} not supported (float32 / float64)")
            prep = lambda v: None if v is None else \
                torch.as_tensor(v, dtype=d.dtype, device=self._cdev).broadcast_to(out_shape).contiguous()
            bb, mm = prep(b), prep(mask)
            fused = self._stored is None and self._packed is None and ch_axis is None and \
                (n_vol == 1 or rpv > 0) and self._rays.n > 0
            if fused:
                n = self._rays.n
                out = torch.empty(n, dtype=d.dtype, device=self._cdev)
                grad = torch.zeros((n_vol, self._tables.nr, self._tables.ne, self._tables.na),
                                   dtype=d.dtype, device=self._cdev)
                scratch = self._scratch()
                ptr = lambda v: None if v is None else v.data_ptr()
                self._plan_cache_begin()
                with torch.cuda.device(self._cdev):
                    rc = self._lib.sphrt_normal(self._tables.c, self._rays.c, d.data_ptr(), _DTYPES[d.dtype],
                                                n_vol, rpv, out.data_ptr(), ptr(bb), ptr(mm), float(scale),
                                                grad.data_ptr(), scratch.data_ptr(), scratch.numel(),
                                                self._stream())
                self._plan_cache_end(ok=rc == 0)
                if rc == 0:
                    return out.reshape(out_shape).to(density.device), grad.reshape(shape).to(density.device)
                if rc != -1:                               # -1: not served by the warp walker
                    _cabi.check(rc)
            y = self._apply_forward(d)
            w = y if bb is None else y - bb
            if mm is not None:
                w = w * mm
            g = self._apply_adjoint(w * scale, shape)
            return y.to(density.device), g.to(density.device)

    def forward(self, density, out=None):
        """``op(density)`` without autograd and with an optional preallocated result.

        ``density`` / ``out`` may live on the host; pinned host buffers are copied
        asynchronously on the current stream and the call returns after the result has
        landed in ``out`` (this is the end-to-end path timed by bench.py)."""
        with torch.no_grad():
            host_in = density.device.type == "cpu"
            d = density.to(self._cdev, non_blocking=True) if host_in else density
            y = self._apply_forward(d)
            if out is None:
                return y.to(density.device)
            out.copy_(y.reshape(out.shape), non_blocking=True)
            if out.device.type == "cpu":
                torch.cuda.current_stream(self._cdev).synchronize()
            return out

    def adjoint(self, line_integrations, out=None, density_shape=None):
        """Back-projection ``A^T y`` without autograd; ``density_shape`` defaults to the
        grid shape (give ``(C, N_r, N_e, N_a)`` for a multi-channel ``y``)."""
        with torch.no_grad():
            t = self._tables
            shape = tuple(density_shape) if density_shape is not None else tuple(self.grid.shape)
            y = line_integrations
            yd = y.to(self._cdev, non_blocking=True) if y.device.type == "cpu" else y
            g = self._apply_adjoint(yd, shape)
            if out is None:
                return g.to(y.device)
            out.copy_(g.reshape(out.shape), non_blocking=True)
            if out.device.type == "cpu":
                torch.cuda.current_stream(self._cdev).synchronize()
            return out

    def __repr__(self):
        if self.dynamic:
            return f"Operator({(self.geom.shape[0], *self.grid.shape)} → {self.geom.shape})"
        return f"Operator({self.grid.shape} → {self.geom.shape})"

    # ---- parity / metric / debug exports ---------------------------------------------------
    def segment_counts(self):
        """Per-ray number of canonical segments (non-zero length, consecutive identical
        voxels merged) and their total -- the unit of the ray-segments/s metric."""
        count = torch.empty(self._rays.n, dtype=torch.int32, device=self._cdev)
        total = torch.zeros(1, dtype=torch.int64, device=self._cdev)
        scratch = self._scratch()
        with torch.cuda.device(self._cdev):
            _cabi.check(self._lib.sphrt_count_segments(
                self._tables.c, self._rays.c, count.data_ptr(), total.data_ptr(),
                scratch.data_ptr(), scratch.numel(), self._stream()))
        return count.reshape(self._rays.shape), int(total.item())

    def segments(self, max_seg=None):
        """Canonical form of the reference's ``regs`` / ``lens``: ``vox (*rays, max_seg, 3)``
        int32 (r, e, a), ``lens (*rays, max_seg)`` float64 and ``count (*rays)``."""
        t = self._tables
        if max_seg is None:
            max_seg = 2 * (t.nr + 1) + 2 * (t.ne + 1) + (t.na + 1) + 1
        n = self._rays.n
        vox = torch.full((n, max_seg, 3), -9, dtype=torch.int32, device=self._cdev)
        lens = torch.zeros((n, max_seg), dtype=torch.float64, device=self._cdev)
        count = torch.zeros(n, dtype=torch.int32, device=self._cdev)
        scratch = self._scratch()
        with torch.cuda.device(self._cdev):
            _cabi.check(self._lib.sphrt_trace_segments(
                self._tables.c, self._rays.c, max_seg, vox.data_ptr(), lens.data_ptr(),
                count.data_ptr(), scratch.data_ptr(), scratch.numel(), self._stream()))
        G = self._rays.shape
        return vox.reshape(G + (max_seg, 3)), lens.reshape(G + (max_seg,)), count.reshape(G)

    def walker(self):
        """'cooperative' or 'per-thread': the kernel that walks this operator's ordinary rays."""
        import os
        if self._rays.c.walker and not os.environ.get("SPHRT_WALKER"):
            return "cooperative" if self._rays.c.walker == 1 else "per-thread"
        kind = int(self._lib.sphrt_walker_kind(self._tables.c, self._rays.n))
        _cabi.check(min(kind, 0))
        return "cooperative" if kind == 1 else "per-thread"

    def events(self, max_ev=None):
        """Raw export of the warp-cooperative walker (parity tests): per ray every event in
        merged order as ``vox (n, max_ev, 3)``, ``lens (n, max_ev)``, ``count (n)`` (-1 for
        rays that walker hands to the per-thread walker) and the walker's statistics
        ``(cooperative, listed, handed over, failed verification)``."""
        t = self._tables
        if max_ev is None:
            max_ev = 2 * (t.nr + 1) + 2 * (t.ne + 1) + (t.na + 1) + 1
        n = self._rays.n
        vox = torch.full((n, max_ev, 3), -9, dtype=torch.int32, device=self._cdev)
        lens = torch.zeros((n, max_ev), dtype=torch.float64, device=self._cdev)
        count = torch.full((n,), -1, dtype=torch.int32, device=self._cdev)
        stats = torch.zeros(4, dtype=torch.int64, device=self._cdev)
        scratch = self._scratch()
        with torch.cuda.device(self._cdev):
            _cabi.check(self._lib.sphrt_trace_events(
                self._tables.c, self._rays.c, max_ev, vox.data_ptr(), lens.data_ptr(),
                count.data_ptr(), stats.data_ptr(), scratch.data_ptr(), scratch.numel(),
                self._stream()))
        return vox, lens, count, tuple(stats.tolist())

    def _print_los(self, los):
        """debug=True: the slot table of one line of sight in the reference's format (typ / reg /
        intlen / dist / ind / neg per sorted candidate, raytracer.py:175-225).  Display only: the
        candidates come from the CUDA extension (sphrt_candidates: the r_torch / e_torch / a_torch
        tables), the one ray's sort + forward fill + diff (:131-170) is repeated on the host."""
        r, t = self._rays, self._tables
        los = tuple(los) if los is not None else (0,) * len(r.shape)
        i = 0
        for k, n in zip(los, r.shape):
            i = i * n + int(k)
        x = r.starts[i // r.rays_per_start].reshape(1, 3)
        d = r.directions(self._lib, self._stream()).reshape(-1, 3)[i].reshape(1, 3)
        one = _Rays(types.SimpleNamespace(ray_starts=x.cpu(), rays=d.cpu()), self._cdev, t)
        k1 = 2 * (t.nr + 1) + 2 * (t.ne + 1) + (t.na + 1)
        ts = torch.empty((1, k1), dtype=torch.float64, device=self._cdev)
        reg = torch.empty((1, k1), dtype=torch.int32, device=self._cdev)
        geo = torch.empty((1, k1), dtype=torch.int32, device=self._cdev)
        with torch.cuda.device(self._cdev):
            _cabi.check(self._lib.sphrt_candidates(t.c, one.c, ts.data_ptr(), reg.data_ptr(), geo.data_ptr(),
                                                   self._stream()))
        ts, reg, geo = ts[0].cpu(), reg[0].cpu().long(), geo[0].cpu().long()
        nb = (t.nr + 1, t.ne + 1, t.na + 1)
        kind = torch.cat([torch.full((2 * nb[0],), 0), torch.full((2 * nb[1],), 1), torch.full((nb[2],), 2),
                          torch.tensor([-1])])
        ind = torch.cat([torch.arange(nb[0]).repeat(2), torch.arange(nb[1]).repeat(2), torch.arange(nb[2]),
                         torch.tensor([-1])])
        neg = torch.cat([ind[:-1] - geo, torch.tensor([-1])])
        start = one.cells[0, 3:6].cpu().long()                        # find_starts of the start point
        regs = torch.full((3, k1 + 1), -2, dtype=torch.long)
        for f in range(3):
            regs[f, :k1][kind[:-1] == f] = reg[kind[:-1] == f]
        regs[:, k1] = start
        ts = torch.cat([ts, torch.zeros(1, dtype=torch.float64)])
        ts[ts == float("inf")] = float("nan")                       # the table marks "no root" with +inf
        regs[:, ts < 0] = -2                                          # behind the start (:127)
        ts_s, order = ts.sort()
        regs = regs[:, order]
        state = start.clone()
        for k in range(k1 + 1):                                       # forward fill (:17-45)
            upd = regs[:, k] != -2
            state = torch.where(upd, regs[:, k], state)
            regs[:, k] = state
        lens = torch.diff(ts_s, append=torch.tensor([float("inf")], dtype=torch.float64))
        lens[lens.isinf() | lens.isnan()] = 0
        shape = torch.tensor([t.nr, t.ne, t.na])[:, None]
        lens[((regs > shape - 1) | (regs < 0)).any(0)] = 0
        kmap = {-1: "?", 0: "r", 1: "e", 2: "a"}
        print("ray_start:", x[0].cpu())
        print("ray:", d[0].cpu())
        print("typ   reg       intlen     dist      ind  neg")
        print("---------------------------------------------")
        for k in range(k1 + 1):
            o = int(order[k])
            print(f"{kmap[int(kind[o])]:<2}", f"r:[{int(regs[0, k]):>2},{int(regs[1, k]):>2},{int(regs[2, k]):>2}]",
                  f"l:{float(lens[k]):<6.2f}", f"t:{float(ts_s[k]):<10.2f}" f"i:{int(ind[o]):<2}", f"n:{int(neg[o]):<2}")

    def plot(self, *args, **kwargs):
        raise NotImplementedError("plotting is outside the scope of the B200 operator")


def find_starts(grid, xs, ftype=FTYPE, device=DEVICE):
    """Voxel index of start points, ``(3, *xs.shape[:-1])`` int64 with -1 = outside
    (reference raytracer.py:605-644), on ``device`` like the reference's.

    Evaluated with the same host code the Operator bins its start points with
    (:func:`_start_cells`: torch's own cart2sph + searchsorted + fix-ups), because a point that
    sits on a cone or half-plane to the last bit is binned by the last bit of ``atan2``, which
    differs between torch's CPU library and CUDA's; :func:`find_starts_cuda` is the in-kernel
    version the walkers use for entry points."""
    pts = torch.asarray(xs, dtype=torch.float64).cpu()
    lead = tuple(pts.shape[:-1])
    r_b = torch.asarray(grid.r_b, dtype=torch.float64).cpu().reshape(-1)
    e_b = torch.asarray(grid.e_b, dtype=torch.float64).cpu().reshape(-1)
    a_b = torch.asarray(grid.a_b, dtype=torch.float64).cpu().reshape(-1)
    cells = _start_cells(r_b, e_b, a_b, pts.reshape(-1, 3).contiguous())
    out = cells[:, 3:6].to(torch.int64).T.reshape((3,) + lead)
    dev = torch.device(device)
    if dev.type == "cuda" and not torch.cuda.is_available():
        return out
    return out.to(dev)


def find_starts_cuda(grid, xs, device=DEVICE):
    """:func:`find_starts` evaluated by the CUDA extension (``sphrt_find_starts``: CUDA's atan2)."""
    lib = _cabi.load()
    cdev = _compute_device(device)
    tables = _GridTables(grid, cdev)
    pts = torch.asarray(xs, dtype=torch.float64)
    lead = tuple(pts.shape[:-1])
    flat = pts.reshape(-1, 3).contiguous().to(cdev)
    out = torch.empty((3, flat.shape[0]), dtype=torch.int32, device=cdev)
    with torch.cuda.device(cdev):
        _cabi.check(lib.sphrt_find_starts(tables.c, flat.shape[0], flat.data_ptr(), out.data_ptr(),
                                          torch.cuda.current_stream(cdev).cuda_stream))
    return out.to(torch.int64).reshape((3,) + lead).to(pts.device)


def _candidates(family, bounds, xs, rays, device):
    """Candidate table of one boundary family for rays ``xs + t * rays`` (the CUDA extension's
    ``sphrt_candidates`` on a grid whose other two families are dummies)."""
    import types
    lib = _cabi.load()
    cdev = _compute_device(device)
    b = torch.asarray(bounds, dtype=torch.float64).cpu().reshape(-1)
    if len(b) == 1:
        # one boundary = no voxel: evaluate next to a dummy second boundary and keep the first's
        # columns; with a single boundary the reference maps region 0 to "outside" (:321, 464, 532)
        t, reg, inds, neg = _candidates(family, torch.cat((b, b + 1.0)), xs, rays, device)
        cols = [0] if family == "a" else [0, 2]
        reg = reg[..., cols]
        return t[..., cols], torch.where(reg == 0, -1, reg), inds[..., cols], neg[..., cols]
    dummy = torch.tensor([0.0, 1.0], dtype=torch.float64)
    grid = types.SimpleNamespace(r_b=b if family == "r" else dummy, e_b=b if family == "e" else dummy,
                                 a_b=b if family == "a" else dummy)
    tables = _GridTables(grid, cdev)
    xs = torch.asarray(xs, dtype=torch.float64)
    rays = torch.asarray(rays, dtype=torch.float64)
    shape = torch.broadcast_shapes(xs.shape, rays.shape)
    lead = tuple(shape[:-1])
    geom = types.SimpleNamespace(ray_starts=xs.broadcast_to(shape).reshape(-1, 3),
                                 rays=rays.broadcast_to(shape).reshape(-1, 3))
    r = _Rays(geom, cdev, tables)
    k1 = 2 * (tables.nr + 1) + 2 * (tables.ne + 1) + (tables.na + 1)
    t = torch.empty((r.n, k1), dtype=torch.float64, device=cdev)
    reg = torch.empty((r.n, k1), dtype=torch.int32, device=cdev)
    geo = torch.empty((r.n, k1), dtype=torch.int32, device=cdev)
    with torch.cuda.device(cdev):
        _cabi.check(lib.sphrt_candidates(tables.c, r.c, t.data_ptr(), reg.data_ptr(), geo.data_ptr(),
                                         torch.cuda.current_stream(cdev).cuda_stream))
    n1 = len(b)
    o = {"r": 0, "e": 2 * (tables.nr + 1), "a": 2 * (tables.nr + 1) + 2 * (tables.ne + 1)}[family]
    width = n1 if family == "a" else 2 * n1
    sl = slice(o, o + width)
    inds = torch.arange(n1, device=cdev).repeat(1 if family == "a" else 2)
    t, reg, geo = t[:, sl], reg[:, sl].to(torch.int64), geo[:, sl].to(torch.int64)
    neg = inds[None, :] - geo
    out_dev = xs.device
    return (t.reshape(lead + (width,)).to(out_dev), reg.reshape(lead + (width,)).to(out_dev),
            inds.broadcast_to((r.n, width)).reshape(lead + (width,)).to(out_dev),
            neg.reshape(lead + (width,)).to(out_dev))


def r_torch(r, xs, rays, ftype=FTYPE, itype=ITYPE, device=DEVICE):
    """Intersections with the concentric shells ``r`` (reference raytracer.py:248-325):
    ``(t, regions, inds, negative_crossing)``, each ``(*num_rays, 2 * len(r))`` -- first roots
    then second roots.  Evaluated by the CUDA extension with the arithmetic the operator uses
    (the rays are normalised as ``trace_indices`` leaves them, Appendix B of SURVEY.md)."""
    return _candidates("r", r, xs, rays, device)


def e_torch(e, xs, rays, ftype=FTYPE, itype=ITYPE, device=DEVICE):
    """Intersections with the elevation cones ``e`` (reference raytracer.py:328-468), same
    layout as :func:`r_torch`; region -2 marks a glancing crossing."""
    return _candidates("e", e, xs, rays, device)


def a_torch(a_b, xs, rays, ftype=FTYPE, itype=ITYPE, device=DEVICE):
    """Intersections with the azimuth half-planes ``a_b`` (reference raytracer.py:471-552):
    ``(t, regions, inds, negative_crossing)``, each ``(*num_rays, len(a_b))``."""
    return _candidates("a", a_b, xs, rays, device)
