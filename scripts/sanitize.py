"""A small pass over every kernel family of the library for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck  python scripts/sanitize.py
    compute-sanitizer --tool racecheck python scripts/sanitize.py --lattice-only
Sizes are ShapeNet-like so that the run stays within minutes under the tool's slowdown."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from lattice_net_b200 import Lattice, ModelParams, lattice as L  # noqa: E402
from lattice_net_b200.losses import segmentation_loss  # noqa: E402
from lattice_net_b200.models import LNN  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
pos_np, labels_np = bench.synthetic_cloud(3)
pos, labels = torch.from_numpy(pos_np).to(dev), torch.from_numpy(labels_np).to(dev)
# hash insert / splat / coarsen / neighbour tables
lat = Lattice(bench.CAPACITY, [(bench.SIGMA, 3)])
lat.begin_splat()
idx, w = lat.splat_standalone(pos, torch.randn((bench.NR_POINTS, 8), device=dev))
nv = lat.nr_lattice_vertices()
coarse = lat.create_coarse_verts()
coarse2 = lat.create_coarse_verts_naive(pos)
h = lat.clone_lattice()
h.set_values(torch.randn((nv, 64), device=dev))
fb = torch.randn((9 * 64, 64), device=dev) * 0.05
out = h.convolve_im2row_standalone(fb, 1, h, False, bias=torch.randn(64, device=dev), residual=torch.randn((nv, 64), device=dev))
gi, gf = h.clone_lattice().conv_backward(h, torch.randn((nv, 64), device=dev), fb, 1)
cv = coarse2.convolve_im2row_standalone(fb, 1, h, False)
s = out.slice_standalone_with_precomputation(pos, idx, w)
torch.cuda.synchronize()
print("lattice ops ok: nv", nv, "coarse", coarse.nr_lattice_vertices(), coarse2.nr_lattice_vertices())
if "--lattice-only" not in sys.argv:
    # the whole model, forward + backward, eager (every fused kernel of the step)
    model = LNN(bench.NR_CLASSES, ModelParams(), device=dev)
    lattice = Lattice(bench.CAPACITY, [(bench.SIGMA, 3)])
    logsm, _ = model(lattice, pos, torch.zeros((bench.NR_POINTS, 1), device=dev))
    loss = segmentation_loss(logsm, labels)
    loss.backward()
    torch.cuda.synchronize()
    print("model step ok: loss", float(loss))
