"""One scene-sized convolution forward + weight gradient (10^6 points, nv ~ 4.6e5) per width for an `ncu --set full` capture.
usage: ncu --set full -k regex:"conv_tc3|conv_wgrad_tc" -c 4 python scripts/ncu_big_conv.py [V ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_ops  # noqa: E402
from lattice_net_b200 import Lattice, lattice as L  # noqa: E402

dev = torch.device("cuda", 0)
n, d = 1000000, 3
pos = torch.from_numpy(bench_ops.uniform_cloud(n, d, 0)).to(dev)
lat = Lattice(4 * n, [(0.01351, d)])
lat.begin_splat()
lat.just_create_verts(pos, False)
nv = lat.nr_lattice_vertices()
for V in [int(a) for a in sys.argv[1:]] or [64]:
    h = lat.clone_lattice()
    h.set_values(torch.randn((nv, V), device=dev))
    fb = torch.randn((9 * V, V), device=dev) * 0.05
    L.prepare_filters([(fb, 9, V, V, False)])
    g = torch.randn((nv, V), device=dev)
    for _ in range(2):
        h.convolve_im2row_standalone(fb, 1, h, False)
        h.clone_lattice().conv_weight_grad(h, g, 9, 1)
torch.cuda.synchronize()
print("nv", nv)
