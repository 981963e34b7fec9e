"""Run-to-run reproducibility of the LNN gradients (same weights, same cloud, vertex-0 quirk off so the model is
invariant to vertex numbering).  Prints, per configuration, the worst per-tensor deviation from run 0.
Development aid, run on the GPU box."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from lattice_net_b200 import Lattice, ModelParams, lattice as lm, lattice_modules
from lattice_net_b200.losses import segmentation_loss
from lattice_net_b200.models import LNN
from oracle import cases

dev = torch.device("cuda", 0)
torch.manual_seed(1)
pos = torch.from_numpy(cases.box_surface(2048, 1)).to(dev)
vals = torch.zeros((2048, 1), device=dev)
labels = torch.from_numpy(np.random.RandomState(1).randint(0, 7, 2048)).to(dev)
lattice_modules.REFERENCE_VERTEX0_QUIRK = False
Lattice(60000, [(0.05, 3)])
model = LNN(7, ModelParams(), device=dev)
with torch.no_grad():
    model(Lattice(60000, [(0.05, 3)]), pos, vals)


def run():
    lat = Lattice(60000, [(0.05, 3)])
    for p in model.parameters():
        p.grad = None
    logsm, logits = model(lat, pos, vals)
    loss = torch.nn.functional.nll_loss(logsm, labels) if os.environ.get('LOSS') == 'nll' else segmentation_loss(logsm, labels)
    loss.backward()
    torch.cuda.synchronize()
    return loss.item(), {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}, logits.detach().clone()


def compare(tag, reps=6):
    l0, g0, lg0 = run()
    worst = (0.0, "")
    worst_logit = 0.0
    bad = {}
    for r in range(1, reps):
        l, g, lg = run()
        worst_logit = max(worst_logit, float((lg - lg0).abs().max() / lg0.abs().max()))
        for n in g0:
            e = float((g[n] - g0[n]).abs().max() / g0[n].abs().max().clamp(min=1e-30))
            if e > worst[0]:
                worst = (e, n)
            if e > 1e-3:
                bad[n] = max(bad.get(n, 0.0), e)
    print(f"[{tag}] loss {l0:.6f}  worst logit dev {worst_logit:.2e}  worst grad dev {worst[0]:.3e} ({worst[1]})  tensors > 1e-3: {len(bad)}")
    for n, e in sorted(bad.items(), key=lambda kv: -kv[1])[:8]:
        print(f"      {e:.3e}  {n}")


from lattice_net_b200.graphed import estimate_vertex_bounds

clouds = [torch.from_numpy(cases.box_surface(2048, sd)).to(dev) for sd in range(6)]
bounds = estimate_vertex_bounds(60000, [(0.05, 3)], clouds, 4)
print("bounds", bounds)
lm.set_conv_precision(int(os.environ.get("PREC", "0")))
for sd in range(6):
    pos = clouds[sd]
    labels = torch.from_numpy(np.random.RandomState(sd).randint(0, 7, 2048)).to(dev)
    compare(f"cloud {sd} dynamic", reps=4)
    l0, g0, lg0 = run()

    def run_static():
        lat = Lattice(60000, [(0.05, 3)])
        lat.set_vertex_bounds(bounds)
        for p in model.parameters():
            p.grad = None
        logsm, logits = model(lat, pos, vals)
        loss = torch.nn.functional.nll_loss(logsm, labels) if os.environ.get('LOSS') == 'nll' else segmentation_loss(logsm, labels)
        loss.backward()
        torch.cuda.synchronize()
        return loss.item(), {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}

    for r in range(3):
        l1, g1 = run_static()
        errs = sorted(((float((g1[n] - g0[n]).abs().max() / g0[n].abs().max().clamp(min=1e-30)), n) for n in g0), reverse=True)
        print(f"   static-vs-dynamic run {r}: loss {l0:.6f} vs {l1:.6f}; worst {errs[0][0]:.3e} {errs[0][1]}; >1e-3: {sum(e > 1e-3 for e, _ in errs)}")
        for e, n in errs[:4]:
            if e > 1e-3:
                print(f"        {e:.3e} {n}")
