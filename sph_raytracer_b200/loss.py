"""Loss functions for tomographic retrievals (mirror of reference ``sph_raytracer/loss.py``).

A loss is called as ``loss(f, y, d, c)`` with the forward operator, the measurements, the
current density and the model coefficients; ``3 * SquareLoss()`` sets its weight.
"""
import torch

__all__ = ["Loss", "SquareLoss", "FusedSquareLoss", "SquareRelLoss", "AbsLoss", "CheaterLoss",
           "NegRegularizer", "NegSumRegularizer"]


class Loss:
    """Weighted loss term.

    Args:
        projection_mask: weights / mask applied to the line integrals
        volume_mask: mask applied to the density before projecting
        lam (float): weight of this term
        use_grad (bool): False evaluates under ``torch.no_grad`` (monitoring only)
    """

    kind = "regularizer"

    def __init__(self, *args, projection_mask=1, volume_mask=1, lam=1, use_grad=True, **kwargs):
        self.projection_mask, self.volume_mask = projection_mask, volume_mask
        self.lam, self.use_grad = lam, use_grad

    def compute(self, f, y, d, c):
        raise NotImplementedError

    def __call__(self, f, y, d, c):
        if self.use_grad:
            value = self.compute(f, y, d, c)
        else:
            with torch.no_grad():
                value = self.compute(f, y, d, c)
        return None if value is None else self.lam * value

    def __mul__(self, weight):
        self.lam = weight
        return self

    __rmul__ = __mul__

    def __repr__(self):
        return f"{self.lam:.0e} * {type(self).__name__}"

    def _residual(self, f, y, d):
        return y - f(d * self.volume_mask)


class SquareLoss(Loss):
    """Mean squared residual of the line integrals."""
    kind = "fidelity"

    def compute(self, f, y, d, c):
        return torch.mean(self.projection_mask * self._residual(f, y, d) ** 2)


class _FusedSquare(torch.autograd.Function):
    """mean(mask * (y - A d)^2) whose forward pass also produces its gradient A^T (2 mask (A d - y) / N)
    in the same walk of the rays (``Operator.normal``)."""

    @staticmethod
    def forward(ctx, d, op, y, mask):
        yhat, g = op.normal(d, b=y, mask=mask, scale=2.0 / max(y.numel(), 1))
        ctx.save_for_backward(g)
        res = (y - yhat) ** 2
        if mask is not None:
            res = res * mask
        return res.mean()

    @staticmethod
    def backward(ctx, grad_out):
        (g,) = ctx.saved_tensors
        return g * grad_out, None, None, None


class FusedSquareLoss(Loss):
    """:class:`SquareLoss` (same value, same gradient) for an operator that has ``normal()``: the
    forward projection of an iteration and the back-projection of its residual share ONE walk of
    the rays instead of two.  Drop-in for ``SquareLoss`` in ``gd`` (also the reference's own)."""
    kind = "fidelity"

    def compute(self, f, y, d, c):
        if not hasattr(f, "normal"):
            return torch.mean(self.projection_mask * self._residual(f, y, d) ** 2)
        pm = self.projection_mask
        mask = None if isinstance(pm, (int, float)) and pm == 1 else torch.as_tensor(pm, dtype=y.dtype, device=y.device)
        return _FusedSquare.apply(d * self.volume_mask, f, y, mask)


class SquareRelLoss(Loss):
    """Mean squared relative residual (pixels with y == 0 do not contribute)."""
    kind = "fidelity"

    def compute(self, f, y, d, c):
        res = self._residual(f, y, d)
        nz = y != 0
        rel = torch.zeros_like(y)
        rel[nz] = res[nz] / y[nz]
        return torch.mean((self.projection_mask * rel) ** 2)


class AbsLoss(Loss):
    """Mean absolute residual."""
    kind = "fidelity"

    def compute(self, f, y, d, c):
        return torch.mean(self.projection_mask * self._residual(f, y, d).abs())


class CheaterLoss(Loss):
    """Squared distance to a known ground-truth density (monitoring, ``kind = 'oracle'``)."""
    kind = "oracle"

    def __init__(self, density_truth, *args, **kwargs):
        self.density_truth = density_truth
        super().__init__(**kwargs)

    def compute(self, f, y, d, c):
        return torch.mean(self.volume_mask * (d - self.density_truth) ** 2)


class NegRegularizer(Loss):
    """Mean magnitude of the negative voxels."""

    def compute(self, f, y, d, c):
        return torch.mean(torch.abs(self.volume_mask * d.clip(max=0)))


class NegSumRegularizer(Loss):
    """Summed magnitude of the negative voxels."""

    def compute(self, f, y, d, c):
        return torch.sum(torch.abs(self.volume_mask * d.clip(max=0)))
