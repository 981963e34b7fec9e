"""GPU-vs-CPU-port gradient agreement of the LatticeNet step per cloud and attempt (which statistics are robust to
ReLU-gate flips between two fp32 evaluations?).  Development aid for tests/test_gpu_parity.py::test_lnn_model_matches_cpu_port."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from lattice_net_b200 import Lattice, ModelParams
from lattice_net_b200.models import LNN
from oracle import cases, cpu_port

dev = torch.device("cuda", 0)
torch.manual_seed(0)
Lattice(60000, [(0.05, 3)])        # sets the static expected position dimension the module constructors read
model = LNN(7, ModelParams(), device=dev)
cpu = None
for seed in (0, 3, 4, 1):
    pos_np = cases.box_surface(2048, seed)
    labels_np = np.random.RandomState(3).randint(0, 7, 2048)
    pos, vals, labels = torch.from_numpy(pos_np).to(dev), torch.zeros((2048, 1), device=dev), torch.from_numpy(labels_np).to(dev)
    for attempt in range(4):
        lattice = Lattice(60000, [(0.05, 3)])
        for p in model.parameters():
            p.grad = None
        logsm, logits = model(lattice, pos, vals)
        loss = torch.nn.functional.nll_loss(logsm, labels)
        loss.backward()
        l1 = model.last_level1_lattice
        keys = l1.hash_table().m_keys_tensor[:l1.nr_lattice_vertices()].cpu().numpy()
        if cpu is None:
            cpu = cpu_port.CpuLNN(7, ModelParams())
            cpu.load_state_dict({k: v.detach().cpu() for k, v in model.state_dict().items()}, strict=True)
        for p in cpu.parameters():
            p.grad = None
        clogsm, clogits = cpu(pos_np, torch.zeros(2048, 1), [0.05] * 3, level1_keys=keys)
        closs = torch.nn.functional.nll_loss(clogsm, torch.from_numpy(labels_np))
        closs.backward()
        cg_all = dict(cpu.named_parameters())
        worst, rel, fa, fb, over = (0.0, ""), [], [], [], 0
        num = den = 0.0
        for name, p in model.named_parameters():
            if p.grad is None:
                continue
            g, cg = p.grad.detach().cpu().numpy().astype(np.float64).ravel(), cg_all[name].grad.numpy().astype(np.float64).ravel()
            e = np.abs(g - cg).max() / max(np.abs(cg).max(), 1e-30)
            worst = max(worst, (e, name))
            over += e > 2e-2
            rel.append(np.abs(g - cg) / max(np.abs(cg).max(), 1e-30))
            fa.append(g); fb.append(cg)
            num += ((g - cg) ** 2).sum(); den += (cg ** 2).sum()
        fa, fb, rel = np.concatenate(fa), np.concatenate(fb), np.concatenate(rel)
        cos = float(fa @ fb / (np.linalg.norm(fa) * np.linalg.norm(fb)))
        ldev = float((logits.detach().cpu() - clogits.detach()).abs().max() / clogits.detach().abs().max())
        print(f"cloud {seed} attempt {attempt}: logits {ldev:.2e} loss {abs(loss.item() - closs.item()) / abs(closs.item()):.1e} "
              f"worst {worst[0]:.3e} ({worst[1][-45:]}) tensors>2e-2: {over} L2 {np.sqrt(num / den):.3e} cos {cos:.6f} "
              f"median {np.median(rel):.2e} p99 {np.quantile(rel, 0.99):.2e}", flush=True)
