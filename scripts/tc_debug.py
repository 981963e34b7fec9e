"""First-light check of the tcgen05 conv kernel: prints error statistics instead of asserting."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lattice_net_b200 import Lattice, lattice as lm
from oracle import cases
torch.manual_seed(0)
pos = torch.from_numpy(cases.box_surface(2048, 0)).cuda()
lat = Lattice(60000, [(0.05, 3)])
lat.begin_splat(); lat.splat_standalone(pos, torch.zeros((2048, 1), device="cuda"))
nv = lat.nr_lattice_vertices()
for cin, cout in ((32, 32), (64, 128), (128, 256), (96, 16)):
    x = torch.randn((nv, cin), device="cuda")
    w = torch.randn((9 * cin, cout), device="cuda") * 0.1
    l2 = lat.clone_lattice(); l2.set_values(x)
    lm.set_conv_precision(0)
    ref = l2.convolve_im2row_standalone(w, 1, l2, False).values()
    for prec in (2, 1):
        lm.set_conv_precision(prec)
        try:
            out = l2.convolve_im2row_standalone(w, 1, l2, False).values()
            torch.cuda.synchronize()
            err = (out - ref).abs().max().item() / ref.abs().max().item()
            print(f"cin={cin} cout={cout} prec={prec}: max rel err {err:.3e}  nan={torch.isnan(out).any().item()}", flush=True)
        except Exception as e:
            print(f"cin={cin} cout={cout} prec={prec}: FAILED {e}", flush=True)
            raise
lm.set_conv_precision(0)
