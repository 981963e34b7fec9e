"""A handful of ShapeNet-sized convolutions for an `ncu --set full` capture (why does 128->128 at 1408 rows take 2x the
32->32 one?).  usage: ncu --set full -k regex:conv_tc2 -c 6 python scripts/ncu_small_conv.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from lattice_net_b200 import Lattice, lattice as L  # noqa: E402

dev = torch.device("cuda", 0)
pos = torch.from_numpy(bench.synthetic_cloud(7)[0]).to(dev)
lat = Lattice(bench.CAPACITY, [(bench.SIGMA, 3)])
lat.set_vertex_bounds([1408, 384, 128, 128])
lat.begin_splat()
lat.just_create_verts(pos, False)
arena = L.ZeroArena(64 << 20, dev)
L.set_zero_arena(arena)
for cin, cout in [(32, 32), (128, 128), (128, 32)]:
    h = lat.clone_lattice()
    nv = h.nr_lattice_vertices()
    x = torch.randn((nv, cin), device=dev)
    fb = torch.randn((9 * cin, cout), device=dev) * 0.05
    L.prepare_filters([(fb, 9, cin, cout, False)])
    h.set_values(x)
    for _ in range(2):
        arena.reset()
        h.convolve_im2row_standalone(fb, 1, h, False)
torch.cuda.synchronize()
