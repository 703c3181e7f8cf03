"""Multi-GPU use of the operator: rays shard by observation, one process per GPU.

The forward pass needs no communication (every rank integrates its own observations
through a replicated volume); the adjoint scatter-adds into a rank-local gradient volume
and finishes with ONE all-reduce (sum) of that small volume over NCCL / NVLink
(SURVEY.md 8e).  For a dynamic volume sharded by observation the time-slices touched by
different ranks are disjoint, so the same all-reduce simply assembles them (``ShardedOperator``
shards the time axis of a dynamic grid together with the observations).

Works with any initialised ``torch.distributed`` backend (``nccl`` on GPUs; the host logic
is exercised with ``gloo`` in the CPU tests).
"""
import torch
import torch.distributed as dist

from .geometry import ViewGeom, ViewGeomCollection

__all__ = ["shard_bounds", "shard_collection", "shard_detector_rows", "ShardedOperator",
           "allreduce_gradient"]


def shard_bounds(n_items, rank, world):
    """Contiguous block [lo, hi) of ``n_items`` owned by ``rank``; sizes differ by at most 1."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_collection(geoms, rank, world):
    """The observations of a ViewGeomCollection owned by ``rank`` (a new collection)."""
    lo, hi = shard_bounds(len(geoms), rank, world)
    if hi == lo:
        raise ValueError(f"rank {rank} of {world} would own no observation ({len(geoms)} in total)")
    return ViewGeomCollection(*geoms.geoms[lo:hi]), (lo, hi)


def shard_detector_rows(geoms, rank, world):
    """Fewer observations than ranks (a single vantage, ``examples/single_vantage.py``): the
    detector ROWS of all observations, taken together, are sharded instead.

    ``geoms`` is one geometry or a collection; every observation contributes ``shape[-2]`` rows of
    ``shape[-1]`` pixels.  Returns a plain :class:`ViewGeom` with this rank's rows (explicit start
    and direction per pixel, shape ``(rows, W)``) and the block ``(lo, hi)`` of the flattened
    (observation, row) axis it covers, so that the ranks' results concatenate to the full
    ``(n_obs * H, W)`` stack."""
    obs = list(geoms.geoms) if isinstance(geoms, ViewGeomCollection) else [geoms]
    rays = [torch.asarray(g.rays, dtype=torch.float64) for g in obs]
    if any(r.dim() != 3 for r in rays):
        raise ValueError("detector-row sharding needs (H, W) detectors")
    heights = [r.shape[0] for r in rays]
    lo, hi = shard_bounds(sum(heights), rank, world)
    if hi == lo:
        raise ValueError(f"rank {rank} of {world} would own no detector row ({sum(heights)} in total)")
    parts_d, parts_x, first = [], [], 0
    for g, r, h in zip(obs, rays, heights):
        a, b = max(lo, first), min(hi, first + h)
        if a < b:
            d = r[a - first:b - first]
            x = torch.asarray(g.ray_starts, dtype=torch.float64).broadcast_to(r.shape)[a - first:b - first]
            parts_d.append(d); parts_x.append(x)
        first += h
    # (not through ViewGeom.__init__: it normalises the directions once more, and a direction that
    # differs in the last bit is a different ray -- the shards must trace exactly the rays of the
    # unsharded geometry)
    geom = ViewGeom.__new__(ViewGeom)
    geom.ray_starts = torch.cat(parts_x).contiguous()
    geom.rays = torch.cat(parts_d).contiguous()
    geom.shape = geom.rays.shape[:-1]
    return geom, (lo, hi)


def allreduce_gradient(grad, group=None):
    """Sum the rank-local gradient volumes in place (no-op without a process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=group)
    return grad


def _time_slices(grid, lo, hi):
    """A view of a dynamic grid restricted to time-slices [lo, hi) (same spatial boundaries)."""
    import copy
    g = copy.copy(grid)
    shape = (hi - lo, *grid.shape[1:])
    g.shape = type(grid.shape)(*shape) if hasattr(grid.shape, "_fields") else shape
    if getattr(grid, "t", None) is not None:
        g.t = grid.t[lo:hi]
    return g


class _ShardedIntegral(torch.autograd.Function):
    @staticmethod
    def forward(ctx, density, sop):
        ctx.sop = sop
        with torch.enable_grad():
            # dynamic volume: this rank integrates its own observations through their own
            # time-slices (observation o reads slice o, reference raytracer.py:705-706)
            d = (density[sop.lo:sop.hi] if sop.dynamic else density).detach().requires_grad_(True)
            y = sop.local(d)
        ctx.save_for_backward(d, y)
        ctx.full_shape = density.shape
        return y.detach()

    @staticmethod
    def backward(ctx, grad_out):
        d, y = ctx.saved_tensors
        (g,) = torch.autograd.grad(y, d, grad_out)
        if ctx.sop.dynamic:
            # the ranks' time-slices are disjoint: the sum over ranks just assembles them
            full = torch.zeros(ctx.full_shape, dtype=g.dtype, device=g.device)
            full[ctx.sop.lo:ctx.sop.hi] = g
            g = full
        return allreduce_gradient(g.contiguous(), ctx.sop.group), None


class ShardedOperator:
    """Data-parallel operator: this rank traces observations [lo, hi) of ``geoms``.

    ``sop(density)`` returns the line integrals of the local observations; its autograd
    backward returns the *global* gradient (all-reduced), so a replicated optimiser state
    stays identical on every rank.  ``operator_cls`` is the single-GPU operator class.
    """

    def __init__(self, grid, geoms, operator_cls, rank=None, world=None, group=None, **kwargs):
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank, self.world, self.group = rank, world, group
        self.grid = grid
        self.dynamic = bool(getattr(grid, "dynamic", False))
        n_obs = len(geoms.geoms) if isinstance(geoms, ViewGeomCollection) else 1
        self.by_rows = n_obs < world and not self.dynamic
        if self.by_rows:
            # fewer observations than ranks: shard the detector rows (results are (rows, W) blocks
            # of the flattened (observation, row) axis)
            self.geom, (self.lo, self.hi) = shard_detector_rows(geoms, rank, world)
        else:
            self.geom, (self.lo, self.hi) = shard_collection(geoms, rank, world)
        local_grid = grid
        if self.dynamic:
            # the time axis is sharded with the observations: a grid with this rank's slices
            if grid.shape[0] != len(geoms):
                raise IndexError(f"{grid.shape[0]} time-slices cannot be sharded with {len(geoms)} observations")
            local_grid = _time_slices(grid, self.lo, self.hi)
        self.local = operator_cls(local_grid, self.geom, **kwargs)
        self.device = self.local.device

    def __call__(self, density):
        return _ShardedIntegral.apply(density, self)

    def normal(self, density, b=None, mask=None, scale=1.0):
        """Fused normal-equations step of the local observations (``Operator.normal``): the local
        integrals and the GLOBAL (all-reduced) ``A^T (scale * mask * (A x - b))``; ``b`` and ``mask``
        are this rank's blocks.  ``loss.FusedSquareLoss`` finds it through ``hasattr(f, "normal")``."""
        with torch.no_grad():
            d = density[self.lo:self.hi] if self.dynamic else density
            if hasattr(self.local, "normal"):
                y, g = self.local.normal(d, b=b, mask=mask, scale=scale)
            else:
                y = self.local(d)
                w = y if b is None else y - b
                if mask is not None:
                    w = w * mask
                g = self.local.T(w * scale)
            if self.dynamic:
                full = torch.zeros(tuple(density.shape), dtype=g.dtype, device=g.device)
                full[self.lo:self.hi] = g
                g = full
            return y, allreduce_gradient(g.contiguous(), self.group)

    def T(self, line_integrations):
        g = self.local.T(line_integrations)
        if self.dynamic:
            full = torch.zeros(tuple(self.grid.shape), dtype=g.dtype, device=g.device)
            full[self.lo:self.hi] = g
            g = full
        return allreduce_gradient(g, self.group)
