"""Diagnostic: dynamic-eager vs static-eager vs graph replay gradients of one training step."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import cases
from lattice_net_b200 import Lattice, ModelParams
from lattice_net_b200.graphed import GraphedTrainStep, estimate_vertex_bounds
from lattice_net_b200.losses import segmentation_loss
from lattice_net_b200.models import LNN
from lattice_net_b200.parallel import GradBucket

torch.manual_seed(1)
dev = torch.device("cuda", 0)
cuda = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
clouds = [(cuda(cases.box_surface(2048, s)), torch.zeros((2048, 1), device=dev), cuda(np.random.RandomState(s).randint(0, 7, 2048))) for s in (0, 1, 2)]
lat_a = Lattice(60000, [(0.05, 3)])
model_a = LNN(7, ModelParams(), device=dev)
with torch.no_grad():
    model_a(lat_a, *clouds[0][:2])
model_b = copy.deepcopy(model_a)
model_c = copy.deepcopy(model_a)

def grads_of(model, lattice, cloud):
    pos, vals, labels = cloud
    logsm, logits = model(lattice, pos, vals)
    loss = segmentation_loss(logsm, labels)
    for p in model.parameters():
        p.grad = None
    loss.backward()
    l1 = model.last_level1_lattice
    key0 = l1.hash_table().m_keys_tensor[0].cpu().numpy()
    return loss.item(), {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}, key0, logits.detach().clone()

def compare(tag, ga, gb):
    errs = []
    for n in ga:
        a, b = ga[n].double(), gb[n].double()
        errs.append(((a - b).abs().max() / b.abs().max().clamp(min=1e-30)).item())
    names = list(ga)
    worst = np.argsort(errs)[::-1][:5]
    print(tag, "max err %.3e median %.3e" % (max(errs), float(np.median(errs))), [(names[i], "%.2e" % errs[i]) for i in worst])

cloud = clouds[1]
la1, ga1, k1, lg1 = grads_of(model_a, lat_a, cloud)
la2, ga2, k2, lg2 = grads_of(model_a, lat_a, cloud)
print("dynamic eager twice: loss", la1, la2, "key0", k1, k2)
compare("dyn vs dyn", ga1, ga2)
bounds = estimate_vertex_bounds(60000, [(0.05, 3)], [c[0] for c in clouds], 4)
print("bounds", bounds)
lat_c = Lattice(60000, [(0.05, 3)])
lat_c.set_vertex_bounds(bounds)
lc, gc, kc, lgc = grads_of(model_c, lat_c, cloud)
print("static eager: loss", lc, "key0", kc, "nv", [l.nr_lattice_vertices_actual() for l in model_c.last_level_lattices])
compare("static-eager vs dyn", gc, ga1)
print("logits err static-eager vs dyn", ((lgc - lg1).abs().max() / lg1.abs().max()).item())
lat_b = Lattice(60000, [(0.05, 3)])
opt_b = torch.optim.AdamW(model_b.parameters(), lr=1e-3, weight_decay=3e-4, amsgrad=True, fused=True, capturable=True)
bucket_b = GradBucket(model_b.parameters())
step = GraphedTrainStep(model_b, lat_b, opt_b, segmentation_loss, 2048, 3, 1, bounds, bucket_b, example=clouds[0])
lb = step(*cloud)
torch.cuda.synchronize()
gb = {n: p.grad.detach().clone() for n, p in model_b.named_parameters() if p.grad is not None}
print("graph: loss", lb.item(), "nv", step.last_vertex_counts(), "key0", model_b.last_level1_lattice.hash_table().m_keys_tensor[0].cpu().numpy())
compare("graph vs dyn", gb, ga1)
compare("graph vs static-eager", gb, gc)
