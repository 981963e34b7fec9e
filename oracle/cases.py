"""TEST INFRASTRUCTURE -- seeded synthetic inputs shared by the golden-vector generator and the tests
(shapes follow SURVEY.md section 8d)."""
import numpy as np


def box_surface(n, seed, size=(0.8, 0.3, 0.4)):
    """n points uniformly on the faces of an axis-aligned box centred at 0 (ShapeNet-object-like)."""
    rng = np.random.RandomState(seed)
    size = np.asarray(size, np.float64)
    p = (rng.rand(n, 3) - 0.5) * size
    face = rng.randint(0, 3, n)
    side = rng.randint(0, 2, n) * 2 - 1
    p[np.arange(n), face] = 0.5 * size[face] * side
    return p.astype(np.float32)


def near_boundary(n, seed, sigma=0.05):
    """Adversarial cloud: coordinates snapped to a coarse rational grid in *scaled* space, so elevated
    coordinates sit on / within an ulp of simplex faces and remainder-0 ties -- the inputs where an
    arithmetic mismatch (FMA contraction, rsqrt.approx, FTZ) flips a lattice key."""
    rng = np.random.RandomState(seed)
    grid = rng.randint(-40, 41, size=(n, 3)).astype(np.float64) / 8.0          # multiples of 1/8 in scaled units
    jitter = (rng.randint(-2, 3, size=(n, 3)) * 2.0 ** -22) * np.maximum(np.abs(grid), 1.0)
    return ((grid + jitter) * sigma).astype(np.float32)


def cloud_5d(n, seed):
    """xyz + rgb positions for pos_dim = 5 ... (xyz in a box, 2 colour-like channels in [0,1])."""
    rng = np.random.RandomState(seed)
    xyz = (rng.rand(n, 3) - 0.5) * np.array([0.8, 0.6, 0.4])
    col = rng.rand(n, 2)
    return np.concatenate([xyz, col], 1).astype(np.float32)


CASES = {
    # name: (positions fn, sigmas, capacity)
    "shapenet": dict(make=lambda: box_surface(2048, 0), sigmas=[0.05, 0.05, 0.05], capacity=60000),
    "boundary": dict(make=lambda: near_boundary(4096, 1), sigmas=[0.05, 0.05, 0.05], capacity=60000),
    "d5": dict(make=lambda: cloud_5d(1024, 2), sigmas=[0.08, 0.08, 0.08, 0.25, 0.25], capacity=60000),
}


def randn(shape, seed):
    return np.random.RandomState(seed).randn(*shape).astype(np.float32)
