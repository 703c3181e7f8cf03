"""Compute part of the reference's ``plotting.preview3d`` (plotting.py:210-297).

A second client of the operator: one detector, many volumes -- the azimuth-rolled copies of
a static volume (or the time-slices of a dynamic one) are traced as channels that share a
single walk of the rays.  Drawing (matplotlib) is out of scope; this returns the image stack.
"""
import math

import torch

from .geometry import ConeRectGeom
from .raytracer import Operator

__all__ = ["preview3d"]


def preview3d(volume, grid, shape=(256, 256), orbit=True, elev=60, azim=0, device="cuda"):
    """Image stack of a camera orbiting the volume.

    volume ``grid.shape`` (+ optional trailing channel axis) ->
    ``(n_frames, *shape)`` (+ channel axis), n_frames = N_t (dynamic) or N_a (static).
    The camera stays put and the volume is rolled by one azimuth bin per frame.
    """
    nd = len(grid.shape)
    if volume.ndim not in (3, 4, 5) or not 0 <= volume.ndim - nd <= 1:
        raise ValueError(f"volume {tuple(volume.shape)} does not match grid {tuple(grid.shape)}")
    frames = grid.shape.t if grid.dynamic else grid.shape.a
    # camera at 4 r_max, offset by 1/8 azimuth bin to avoid looking along a grid plane
    r, e = 4 * grid.size.r[1], math.radians(elev)
    a = 0.125 * 2 * math.pi / grid.shape.a + math.radians(azim)
    pos = (r * math.sin(e) * math.cos(a), r * math.sin(e) * math.sin(a), r * math.cos(e))
    op = Operator(grid, ConeRectGeom(shape, pos=pos, fov=(30, 30)), device=device)

    channels = volume.ndim == nd + 1
    vol = volume if channels else volume[..., None]            # (..., r, e, a, C)
    stack = []
    for k in range(frames):
        v = vol[k] if grid.dynamic else vol
        stack.append(torch.roll(v, k if orbit else 0, dims=-2))
    rot = torch.stack(stack)                                    # (frames, r, e, a, C)
    n_ch = rot.shape[-1]
    if grid.dynamic:
        # the dynamic operator traces slice k with the same rays for every k
        out = torch.stack([op(rot[..., c].contiguous()) for c in range(n_ch)], dim=-1)
    else:
        flat = rot.movedim(-1, 0).reshape(n_ch * frames, *rot.shape[1:4]).contiguous()
        out = op(flat).reshape(n_ch, frames, *shape).movedim(0, -1)
    return out if channels else out[..., 0]
