"""One launch of each hot kernel at a sweep size (BASELINE configs[4]) for `ncu --set full` captures.
    ncu --set full --clock-control none --import-source on -k regex:'ln::' -o gpurun_out/ops python scripts/ncu_ops.py
Development aid, run on the GPU box."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from lattice_net_b200 import Lattice, lattice as lm

n = int(os.environ.get("N", "1000000"))
V = int(os.environ.get("V", "64"))
d = 3
dev = torch.device("cuda", 0)
rng = np.random.RandomState(0)
pos = torch.from_numpy(rng.rand(n, d).astype(np.float32)).to(dev)
sigma = (1.0 / n) ** (1.0 / d) * 2.2
for _ in range(8):                                            # sigma so that nv ~ n/2
    lat = Lattice(4 * n, [(sigma, d)])
    lat.begin_splat()
    lat.just_create_verts(pos, False)
    nv = lat.nr_lattice_vertices()
    if nv > 0.65 * n:
        sigma *= 1.25
    elif nv < 0.35 * n:
        sigma *= 0.85
    else:
        break
lat = Lattice(4 * n, [(sigma, d)])
lat.begin_splat()
x = torch.randn((n, V), device=dev)
idx, w = lat.splat_standalone(pos, x)                       # splat_build + splat_accumulate
nv = lat.nr_lattice_vertices()
print("n", n, "nv", nv, "V", V)
l2 = lat.clone_lattice()
lv = torch.randn((nv, V), device=dev)
l2.set_values(lv)
sl = l2.slice_standalone_with_precomputation(pos, idx, w)    # slice_fwd
l2.slice_backwards_standalone_with_precomputation_no_homogeneous(pos, torch.randn((n, V), device=dev), idx, w)   # slice_bwd
l2.set_values(lv)
F = 9
fb = torch.randn((F * V, V), device=dev) * 0.05
for prec in (1, 2):
    lm.set_conv_precision(prec)
    out = l2.convolve_im2row_standalone(fb, 1, l2, False)    # neighbour_table (first) + filter_prep + conv_fwd_tc
lm.set_conv_precision(1)
g = torch.randn((nv, V), device=dev)
q = l2.clone_lattice()
gi, gf = q.conv_backward(l2, g, fb, 1)                       # dgrad (tc) + wgrad
l3 = lat.clone_lattice()
lv8 = torch.randn((nv, 8), device=dev)
l3.set_values(lv8)
ga = l3.gather_standalone_with_precomputation(pos, idx, w)
l2.set_values(lv)
nc = 20
dw = torch.zeros((n, d + 1), device=dev)
cw = torch.randn((nc, V), device=dev)
cb = torch.zeros((nc,), device=dev)
logits = l2.slice_classify_with_precomputation(pos, dw, cw, cb, nc, idx, w)
gl = torch.randn((n, nc), device=dev)
gv, gdw, gcw, gcb = torch.zeros_like(lv), torch.zeros_like(dw), torch.zeros_like(cw), torch.zeros_like(cb)
l2.slice_classify_backwards_with_precomputation(gl, pos, lv, dw, cw, cb, nc, gv, gdw, gcw, gcb, idx, w)
torch.cuda.synchronize()
print("done")
