"""Differential fuzzing of both GPU walkers against the CPU oracle (tools/fuzz.py): random
grids (linear / log radii, partial elevation and azimuth ranges) and rays of five kinds -- aimed
from outside, starting inside, starting exactly on grid planes / axes, starting inside on grid
planes / cones, axis-parallel.  Canonical sequences exact, integrals to 1e-9 in float64, on every
oracle-stable ray.  This is the test that found the per-thread walkers' lost t ~ 0 crossing for
starts on a half-plane, the displaced roots of cones with a snapped leading coefficient, and the
last-ulp atan2 differences that made the start-cell table (sphrt_rays.start_cells) necessary."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))


@pytest.mark.parametrize("seed", [1, 2])
def test_fuzz_against_oracle(seed):
    import fuzz
    assert fuzz.run(n_grids=5, n=800, seed=seed, verbose=False) == 0
