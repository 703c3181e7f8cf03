"""End-to-end time of the reference's static_retrieval.py loop (BASELINE configs[2]) on the B200
operator: 100 Adam iterations of gd() with [SquareLoss, NegRegularizer] on the (50,50,50) grid,
50 x ConeCircGeom (100,50) observations; segments regenerated vs stored operator."""
import os, sys, time, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import sph_raytracer_b200 as S
from sph_raytracer_b200.model import FullyDenseModel
from sph_raytracer_b200.loss import SquareLoss, NegRegularizer
from sph_raytracer_b200.retrieval import gd

grid = S.SphericalGrid(shape=(50, 50, 50))
geoms = sum(S.ConeCircGeom((100, 50), (5 * torch.cos(t), 5 * torch.sin(t), 1), fov=(0, 45))
            for t in torch.linspace(0, 2 * torch.pi, 50))
x = torch.zeros(tuple(grid.shape), dtype=torch.float64, device="cuda")
x[:, 25:, :25] = 1; x[:, :25, 25:] = 1
# one-time costs of torch itself (lazy CUDA module loading, first Adam step: ~3 s) stay out of the timing
_op = S.Operator(grid, geoms)
gd(_op, _op(x).detach(), FullyDenseModel(grid), lr=1e-1, num_iterations=2, progress_bar=False, device="cuda")
for stored in (False, True):
    t0 = time.perf_counter()
    op = S.Operator(grid, geoms, store_segments=torch.float64 if stored else False)
    torch.cuda.synchronize(); t_build = time.perf_counter() - t0
    torch.cuda.reset_peak_memory_stats()
    y = op(x).detach()
    model = FullyDenseModel(grid)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    coeffs, _, losses = gd(op, y, model, lr=1e-1, num_iterations=100, loss_fns=[SquareLoss(), NegRegularizer()],
                           progress_bar=False, device="cuda")
    torch.cuda.synchronize(); t_loop = time.perf_counter() - t0
    hist = losses[next(iter(losses))]
    print(f"stored={stored}: operator build {t_build * 1e3:7.1f} ms (stored list {op.stored_bytes / 1e6:.0f} MB), "
          f"100 iterations {t_loop * 1e3:7.1f} ms ({t_loop * 10:.2f} ms / iteration), "
          f"loss {float(hist[0]):.4g} -> {float(hist[-1]):.4g}, peak memory {torch.cuda.max_memory_allocated() / 1e6:.0f} MB", flush=True)
