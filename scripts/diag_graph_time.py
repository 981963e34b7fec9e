"""Where does the replay time of the benched step go?  CPU cost of cudaGraphLaunch vs device time, with and without the
L2 flush between replays.  usage (GPU box): python scripts/diag_graph_time.py"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from lattice_net_b200.graphed import GraphedTrainStep, estimate_vertex_bounds
    from lattice_net_b200.losses import segmentation_loss
    from lattice_net_b200.optim import FlatAdamW
    from lattice_net_b200.parallel import GradBucket
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    if "--no-pdl" in sys.argv:
        from lattice_net_b200 import _cabi
        _cabi.load().ln_set_programmatic_launch(0)
        print("programmatic dependent launch OFF")
    lattice, model = bench.build_training(dev)
    clouds = [bench.synthetic_cloud(i) for i in range(8)]
    dc = [(torch.from_numpy(p).to(dev), torch.zeros((bench.NR_POINTS, 1), device=dev), torch.from_numpy(l).to(dev)) for p, l in clouds]
    with torch.no_grad():
        model(lattice, *dc[0][:2])
    bucket = GradBucket(model.parameters())
    opt = FlatAdamW(bucket, lr=1e-3, weight_decay=3e-4)
    bounds = estimate_vertex_bounds(bench.CAPACITY, [(bench.SIGMA, 3)], [c[0] for c in dc], 4)
    step = GraphedTrainStep(model, lattice, opt, segmentation_loss, bench.NR_POINTS, 3, 1, bounds, bucket, example=dc[0])
    g = step.graphs[0]
    flush = torch.empty((256 << 20) // 4, dtype=torch.float32, device=dev)
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    n = 50
    # device time, back to back, warm L2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for _ in range(n):
        g.replay()
    t_cpu = (time.perf_counter() - t0) / n
    e1.record()
    torch.cuda.synchronize()
    print(f"back-to-back replays: device {e0.elapsed_time(e1) / n:.3f} ms/replay, CPU time inside replay() {t_cpu * 1e3:.3f} ms")
    # one replay at a time (CPU launch cost exposed), warm L2
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"isolated replays (warm L2): median {ts[n // 2]:.3f} ms, min {ts[0]:.3f} ms")
    ts = []
    for i in range(n):
        flush.fill_(float(i))
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"isolated replays (L2 flushed): median {ts[n // 2]:.3f} ms, min {ts[0]:.3f} ms")
    print(f"kernels of this library per replay: {step.launches_per_step}")


if __name__ == "__main__":
    main()
