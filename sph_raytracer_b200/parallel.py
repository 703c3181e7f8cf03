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

from .geometry import ViewGeomCollection

__all__ = ["shard_bounds", "shard_collection", "ShardedOperator", "allreduce_gradient"]


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
        self.geom, (self.lo, self.hi) = shard_collection(geoms, rank, world)
        self.dynamic = bool(getattr(grid, "dynamic", False))
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

    def T(self, line_integrations):
        g = self.local.T(line_integrations)
        if self.dynamic:
            full = torch.zeros(tuple(self.grid.shape), dtype=g.dtype, device=g.device)
            full[self.lo:self.hi] = g
            g = full
        return allreduce_gradient(g, self.group)
