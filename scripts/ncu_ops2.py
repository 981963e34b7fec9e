"""One launch of the kernels added late in round 1 (row-tiled GroupNorm, tiled / vectorised slice_classify) at sweep size,
for an `ncu --set full` capture:
    ncu --set full --clock-control none -k regex:'gn_tiled|slice_classify' -c 12 -o gpurun_out/ops2 python scripts/ncu_ops2.py
Development aid, run on the GPU box."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from lattice_net_b200 import Lattice
from lattice_net_b200.lattice_modules import _GroupNormReLU

n, V, nc, d = 1000000, 64, 16, 3
dev = torch.device("cuda", 0)
pos = torch.from_numpy(np.random.RandomState(0).rand(n, d).astype(np.float32)).to(dev)
lat = Lattice(4 * n, [(0.01351, d)])          # sigma of the op sweep: nv ~ n / 2
lat.begin_splat()
idx, w = lat.just_create_verts(pos, True)
nv = lat.nr_lattice_vertices()
print("n", n, "nv", nv)
x = torch.randn((nv, V), device=dev, requires_grad=True)
gamma, beta = torch.ones(V, device=dev), torch.zeros(V, device=dev)
y = _GroupNormReLU.apply(x, gamma, beta, 32, 1e-5, True)
torch.autograd.grad(y, x, torch.randn((nv, V), device=dev))
l2 = lat.clone_lattice()
lv = torch.randn((nv, V), device=dev)
l2.set_values(lv)
dw = torch.zeros((n, d + 1), device=dev)
cw, cb = torch.randn((nc, V), device=dev), torch.zeros((nc,), device=dev)
logits = l2.slice_classify_with_precomputation(pos, dw, cw, cb, nc, idx, w)
gl = torch.randn((n, nc), device=dev)
gv, gdw, gcw, gcb = torch.zeros_like(lv), torch.zeros_like(dw), torch.zeros_like(cw), torch.zeros_like(cb)
l2.slice_classify_backwards_with_precomputation(gl, pos, lv, dw, cw, cb, nc, gv, gdw, gcw, gcb, idx, w)
torch.cuda.synchronize()
print("done")
