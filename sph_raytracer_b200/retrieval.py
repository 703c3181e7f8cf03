"""Iterative tomographic retrieval (mirror of reference ``sph_raytracer/retrieval.py``).

``gd`` is the caller of the hot path: every iteration runs one forward and one adjoint pass
of the operator (reference retrieval.py:90-116).  With a ``parallel.ShardedOperator`` and the
matching shard of the measurements the same loop is data-parallel over GPUs: the backward of
the sharded operator all-reduces the volume gradient, so every rank takes identical steps.
"""
import math

import torch

from .loss import SquareLoss

__all__ = ["gd", "detach_loss"]


def detach_loss(loss):
    """Plain float of a loss value that must not stay in the autograd graph."""
    return float(loss.detach().cpu()) if isinstance(loss, torch.Tensor) else loss


def gd(f, y, model, coeffs=None, num_iterations=100, loss_fns=None, optim=torch.optim.Adam,
       optim_vars=None, progress_bar=True, device=None, **kwargs):
    """Minimise the weighted sum of ``loss_fns`` over the model coefficients.

    Args / returns as the reference's ``gd`` (retrieval.py:24-55): returns
    ``(coeffs, f(model(coeffs)), losses)`` with ``losses[loss_fn]`` the per-iteration history.
    Extra keyword arguments go to the optimiser (e.g. ``lr=1e-1``).
    """
    if loss_fns is None:
        loss_fns = [SquareLoss()]
    if hasattr(f, "grid") and f.grid != model.grid:
        raise ValueError("f and model must have same grid")
    if y is not None:
        y.requires_grad_()
    if coeffs is None:
        coeffs = torch.ones(model.coeffs_shape, requires_grad=True, device=device or f.device,
                            dtype=torch.float64)
    if optim_vars is None:
        optim_vars = [coeffs]
    for var in optim_vars:
        var.requires_grad_()

    try:
        from tqdm import tqdm
    except ImportError:                                   # progress bar is optional
        def tqdm(it, disable=False):
            return it
    optimiser = optim(optim_vars, **kwargs)
    losses = {fn: [] for fn in loss_fns}
    best_loss, best_coeffs = float("inf"), coeffs
    oracle_stat = 0
    bar = tqdm(range(num_iterations), disable=not progress_bar)
    try:
        for _ in bar:
            optimiser.zero_grad()
            density = model(coeffs)
            total = fidelity = regular = 0
            for fn in loss_fns:
                value = fn(f, y, density, coeffs)
                if fn.use_grad and fn.kind != "oracle":
                    total = total + value
                if fn.kind == "oracle" and not math.isnan(value):
                    oracle_stat = value
                elif fn.kind == "fidelity":
                    fidelity = fidelity + value
                elif fn.kind == "regularizer":
                    regular = regular + value
                losses[fn].append(detach_loss(value))
            if hasattr(bar, "set_description"):
                bar.set_description(f"F:{detach_loss(fidelity):.1e} R:{detach_loss(regular):.1e} "
                                    f"O:{detach_loss(oracle_stat) * 100:.0f}")
            if detach_loss(total) < best_loss:
                best_coeffs = coeffs          # like the reference: aliases the live tensor
            total.backward(retain_graph=True)
            optimiser.step()
            if hasattr(model, "proj"):
                coeffs.data = model.proj(coeffs)
    except KeyboardInterrupt:
        pass
    return best_coeffs, f(model(best_coeffs)), losses
