"""Where do the small fill / copy / add kernels of one eager training step come from?  torch.profiler with stacks.
Development aid, run on the GPU box."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

from lattice_net_b200 import Lattice, ModelParams, lattice as lm
from lattice_net_b200.losses import segmentation_loss
from lattice_net_b200.models import LNN
from oracle import cases

dev = torch.device("cuda", 0)
torch.manual_seed(1)
pos = torch.from_numpy(cases.box_surface(2048, 1)).to(dev)
vals = torch.zeros((2048, 1), device=dev)
labels = torch.from_numpy(np.random.RandomState(1).randint(0, 7, 2048)).to(dev)
lm.set_conv_precision(1)
lattice = Lattice(60000, [(0.05, 3)])
model = LNN(7, ModelParams(), device=dev)
with torch.no_grad():
    model(lattice, pos, vals)
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=3e-4, amsgrad=True, fused=True)


def step():
    logsm, _ = model(lattice, pos, vals)
    loss = segmentation_loss(logsm, labels)
    for p in model.parameters():
        p.grad = None
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    step()
    torch.cuda.synchronize()
ka = prof.key_averages(group_by_stack_n=8)
rows = [e for e in ka if e.key in ("aten::fill_", "aten::zero_", "aten::zeros", "aten::zeros_like", "aten::add", "aten::add_", "aten::copy_", "aten::mul", "aten::div", "aten::sum", "aten::contiguous", "aten::clone")]
rows.sort(key=lambda e: -e.count)
for e in rows[:70]:
    stack = [s for s in e.stack if "lattice_net_b200" in s or "losses" in s or "autograd" in s][:3]
    print(f"{e.key:18s} x{e.count:4d}  " + " <- ".join(s.split('/')[-1][:70] for s in stack))
print("--- kernel counts ---")
kc = {}
for e in prof.key_averages():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        kc[e.key[:70]] = kc.get(e.key[:70], 0) + e.count
for k, v in sorted(kc.items(), key=lambda kv: -kv[1])[:40]:
    print(f"{v:5d}  {k}")
