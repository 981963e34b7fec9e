"""The exact ln_conv_fwd call bench.py's `roofline` object times (level-1 lattice convolution 128 -> 128 of the
ShapeNet workload), launched a few times outside a graph so that `ncu --set full` can capture it:
    ncu --set full --clock-control none --import-source on -k regex:conv_tc3 -o gpurun_out/bench_conv python scripts/ncu_bench_conv.py
scripts/roofline_traffic.py turns the raw export into profiles/roofline_traffic.json, which bench.py reports as
`roofline.traffic`.  Development aid, run on the GPU box."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from lattice_net_b200 import Lattice, lattice as L, set_conv_precision

dev = torch.device("cuda", 0)
set_conv_precision(1)
pos = torch.from_numpy(bench.synthetic_cloud(1234)[0]).to(dev)
lat = Lattice(bench.CAPACITY, [(bench.SIGMA, 3)])
lat.begin_splat()
lat.splat_standalone(pos, torch.zeros((bench.NR_POINTS, 1), device=dev))
nv = lat.nr_lattice_vertices()
lv = torch.randn((nv, 128), device=dev)
fb = torch.randn((9 * 128, 128), device=dev) * 0.05
l2 = lat.clone_lattice()
l2.set_values(lv)
L.prepare_filters([(fb, 9, 128, 128, False)])       # as inside the step: slabs prepared once, the kernel alone per call
arena = L.ZeroArena(nv * 128 + 64, dev)
L.set_zero_arena(arena)
for _ in range(4):
    arena.off = 0
    l2.convolve_im2row_standalone(fb, 1, l2, False)
torch.cuda.synchronize()
print("nv", nv)
