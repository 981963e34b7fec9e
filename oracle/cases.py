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


def kitti_like(n, seed):
    """SemanticKITTI-sized lidar-like scan (BASELINE configs[2]): range concentrated near the
    sensor (2 m + exponential, clipped at 60 m), 70 % of the returns on the ground plane z = -1.7 +- 0.05,
    the rest between the ground and 3 m.  With sigma 0.9 this stays well below the cfg's 100 000-slot table."""
    rng = np.random.RandomState(seed)
    r = np.minimum(2.0 + rng.exponential(9.0, n), 60.0)
    a = rng.uniform(0.0, 2.0 * np.pi, n)
    ground = rng.rand(n) < 0.7
    z = np.where(ground, -1.7 + rng.randn(n) * 0.05, rng.uniform(-1.7, 3.0, n))
    return np.stack([r * np.cos(a), r * np.sin(a), z], 1).astype(np.float32)


def scannet_like(n, seed):
    """ScanNet-sized indoor scene (BASELINE configs[3]): points on the faces of an 8 x 6 x 3 m room
    plus random boxes (furniture); values = rgb (3) + height (1)."""
    rng = np.random.RandomState(seed)
    room = np.array([8.0, 6.0, 3.0])
    nb = n // 3
    p = box_surface(n - nb, seed + 1, size=room).astype(np.float64) + room / 2
    boxes = []
    per = nb // 8
    for b in range(8):
        size = rng.uniform(0.4, 1.6, 3) * np.array([1.0, 1.0, 0.6])
        centre = np.array([rng.uniform(1, 7), rng.uniform(1, 5), size[2] / 2])
        m = per if b < 7 else nb - 7 * per
        boxes.append(box_surface(m, seed + 10 + b, size=size).astype(np.float64) + centre)
    p = np.concatenate([p] + boxes, 0)
    p = p[rng.permutation(len(p))]
    rgb = rng.rand(len(p), 3)
    vals = np.concatenate([rgb, p[:, 2:3] / 3.0], 1)
    return p.astype(np.float32), vals.astype(np.float32)
