"""Where does the host time of one LatticeNet training step go?  cProfile over a few steps + a
torch.profiler kernel/launch count.  Development aid, run on the GPU box."""
import cProfile, pstats, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from lattice_net_b200 import set_conv_precision, _cabi
from lattice_net_b200.parallel import GradBucket

set_conv_precision(int(os.environ.get("CONV_PRECISION", "1")))
dev = torch.device("cuda", 0)
lattice, model = bench.build_training(dev)
clouds = [bench.synthetic_cloud(i) for i in range(8)]
dc = [(torch.from_numpy(p).to(dev), torch.zeros((bench.NR_POINTS, 1), device=dev), torch.from_numpy(l).to(dev)) for p, l in clouds]
with torch.no_grad():
    model(lattice, *dc[0][:2])
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=3e-4, amsgrad=True, fused=True)
bucket = GradBucket(model.parameters())
def step(i):
    return bench.train_step(model, lattice, *dc[i % 8], opt, bucket, 1)
for i in range(5):
    step(i)
torch.cuda.synchronize()
t = time.time()
for i in range(20):
    step(i)
torch.cuda.synchronize()
print(f"wall per step: {(time.time() - t) / 20 * 1e3:.2f} ms")
# host-only time (no sync until the end) split by phase
import contextlib
from lattice_net_b200.losses import segmentation_loss
def phase_times(n=20):
    tf = tb = to = 0.0
    for i in range(n):
        pos, vals, labels = dc[i % 8]
        t0 = time.time(); ls, _ = model(lattice, pos, vals); loss = segmentation_loss(ls, labels); t1 = time.time()
        bucket.zero(); loss.backward(); t2 = time.time()
        opt.step(); t3 = time.time()
        tf += t1 - t0; tb += t2 - t1; to += t3 - t2
    torch.cuda.synchronize()
    print(f"host time per step: forward+loss {tf / n * 1e3:.2f} ms, backward {tb / n * 1e3:.2f} ms, optimizer {to / n * 1e3:.2f} ms")
phase_times()
pr = cProfile.Profile()
pr.enable()
for i in range(10):
    step(i)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(35)
print(s.getvalue()[:6000])
from torch.profiler import profile, ProfilerActivity
_cabi.reset_launch_count()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
print("our launches per step:", _cabi.launch_count() / 3)
ev = prof.key_averages()
kern = [e for e in ev if e.device_type.name == "CUDA"] if hasattr(ev[0], "device_type") else []
tot_k = sum(e.count for e in ev if getattr(e, "self_device_time_total", 0) > 0)
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=25, max_name_column_width=60)[:7000])
