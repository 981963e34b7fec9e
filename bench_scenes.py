#!/usr/bin/env python
"""Scene-sized timings for BASELINE.json configs[2] / configs[3] (context numbers beside bench.py's headline line):

  kitti    SemanticKITTI-sized scan, 120 000 points, sigma 0.9, capacity 100 000, 20 classes,
           lnn_train_semantic_kitti.cfg architecture (pointnet start 64 -> levels 64/128/256/512)
  scannet  ScanNet-sized indoor scene, 150 000 points, xyz + rgb + height, sigma 0.08, capacity 5 000 000, 21 classes,
           lnn_train_scannet.cfg architecture (pointnet start 32, blocks [6,6,8] / 8 / [2,2,2])

For each scene: forward + loss + backward of one scan (ms, scans/s, points/s) and the forward-only inference latency,
one CUDA graph per pass on the static-shape lattice (--mode eager: dynamic-shape launches), L2 flushed before every timed
pass, median of --steps passes.

    python bench_scenes.py [--scene kitti|scannet|both] [--impl ours|reference] [--steps 5] [--warmup 2]

`--impl reference` drives the reference's own Python modules and CUDA kernels (oracle/ref_arm.py, as in
`bench.py --impl reference`); run it in its own process.
One JSON line per scene on stdout."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCENES = {
    "kitti": dict(n=120000, sigma=0.9, capacity=100000, nr_classes=20, val_dim=1,
                  model=dict(pointnet_channels_per_layer=[16, 32, 64], pointnet_start_nr_channels=64, nr_downsamples=3,
                             nr_blocks_down_stage=[2, 2, 2], nr_blocks_bottleneck=3, nr_blocks_up_stage=[1, 2, 2],
                             nr_levels_down_with_normal_resnet=3, nr_levels_up_with_normal_resnet=3)),
    "scannet": dict(n=150000, sigma=0.08, capacity=5000000, nr_classes=21, val_dim=4,
                    model=dict(pointnet_channels_per_layer=[16, 32, 64], pointnet_start_nr_channels=32, nr_downsamples=3,
                               nr_blocks_down_stage=[6, 6, 8], nr_blocks_bottleneck=8, nr_blocks_up_stage=[2, 2, 2],
                               nr_levels_down_with_normal_resnet=3, nr_levels_up_with_normal_resnet=3)),
}


def _box_surface(n, rng, size):
    size = np.asarray(size, np.float64)
    p = (rng.rand(n, 3) - 0.5) * size
    face = rng.randint(0, 3, n)
    side = rng.randint(0, 2, n) * 2 - 1
    p[np.arange(n), face] = 0.5 * size[face] * side
    return p


def synthetic_scene(name, seed):
    """Lidar-like scan (range ~ 2 m + Exp(9 m), 70 % ground returns) / indoor room with furniture boxes, rgb + height."""
    rng = np.random.RandomState(seed)
    n = SCENES[name]["n"]
    if name == "kitti":
        r = np.minimum(2.0 + rng.exponential(9.0, n), 60.0)
        a = rng.uniform(0.0, 2.0 * np.pi, n)
        ground = rng.rand(n) < 0.7
        z = np.where(ground, -1.7 + rng.randn(n) * 0.05, rng.uniform(-1.7, 3.0, n))
        return np.stack([r * np.cos(a), r * np.sin(a), z], 1).astype(np.float32), np.zeros((n, 1), np.float32)
    room = np.array([8.0, 6.0, 3.0])
    nb = n // 3
    parts = [_box_surface(n - nb, rng, room) + room / 2]
    per = nb // 8
    for b in range(8):
        size = rng.uniform(0.4, 1.6, 3) * np.array([1.0, 1.0, 0.6])
        centre = np.array([rng.uniform(1, 7), rng.uniform(1, 5), size[2] / 2])
        parts.append(_box_surface(per if b < 7 else nb - 7 * per, rng, size) + centre)
    p = np.concatenate(parts, 0)
    p = p[rng.permutation(len(p))]
    vals = np.concatenate([rng.rand(len(p), 3), p[:, 2:3] / 3.0], 1)
    return p.astype(np.float32), vals.astype(np.float32)


def run_scene(name, impl, steps, warmup, precision, mode="graph"):
    """mode (ours only): "graph" = static-shape lattice + one CUDA graph per training step / per inference pass
    (lattice_net_b200/graphed.py); "eager" = dynamic-shape launches."""
    from lattice_net_b200 import Lattice, ModelParams, set_conv_precision
    from lattice_net_b200.losses import segmentation_loss
    spec = SCENES[name]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    pos_np, vals_np = synthetic_scene(name, 7)
    pos, vals = torch.from_numpy(pos_np).to(dev), torch.from_numpy(vals_np).to(dev)
    labels = torch.from_numpy(np.random.RandomState(1).randint(0, spec["nr_classes"], spec["n"])).to(dev)
    flush = torch.empty((256 << 20) // 4, dtype=torch.float32, device=dev)
    mp = ModelParams(spec["model"])
    if impl == "reference":
        from oracle import ref_arm
        model = ref_arm.build_reference_model(spec["nr_classes"], mp)
        from latticenet_py.lattice.lovasz_loss import LovaszSoftmax
        lattice = ref_arm.RefHandle(spec["capacity"], [spec["sigma"]] * 3)
        lov, nll = LovaszSoftmax(ignore_index=-100), torch.nn.NLLLoss(ignore_index=-100)

        def loss_fn(logsm, lab):
            return 0.5 * lov(logsm, lab) + 0.5 * nll(logsm, lab)
        mode = "eager"
    else:
        from lattice_net_b200.models import LNN
        set_conv_precision(precision)
        lattice = Lattice(spec["capacity"], [(spec["sigma"], 3)])
        model = LNN(spec["nr_classes"], mp, device=dev)
        loss_fn = segmentation_loss

    def train_pass():
        logsm, _ = model(lattice, pos, vals)
        loss = loss_fn(logsm, labels)
        for p in model.parameters():
            p.grad = None
        loss.backward()
        return loss

    def infer_pass():
        with torch.no_grad():
            return model(lattice, pos, vals)[1]

    def timed(fn):
        for _ in range(warmup):
            fn()
        ts = []
        for i in range(steps):
            flush.fill_(float(i))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize(dev)
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    with torch.no_grad():
        model(lattice, pos, vals)          # lazily created parameters
    execution = "eager launches, dynamic-shape lattice"
    if mode == "graph":
        # fwd + loss + bwd (+ a zero-learning-rate AdamW update so that the parameters stay put) as one graph; inference as another
        from lattice_net_b200.graphed import GraphedTrainStep, estimate_vertex_bounds
        from lattice_net_b200.optim import FlatAdamW
        from lattice_net_b200.parallel import GradBucket
        nr_levels = model.nr_downsamples + 1
        bounds = estimate_vertex_bounds(spec["capacity"], [(spec["sigma"], 3)], [pos], nr_levels, headroom=1.1)
        bucket = GradBucket(model.parameters())
        opt = FlatAdamW(bucket, lr=0.0, weight_decay=0.0)
        step = GraphedTrainStep(model, lattice, opt, loss_fn, spec["n"], 3, spec["val_dim"], bounds, bucket, example=(pos, vals, labels), warmup=2)
        fb_ms = timed(lambda: step.graphs[0].replay())
        loss = float(step.loss.item())
        # inference graph on the same static-shape lattice
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            infer_pass()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            logits = infer_pass()
        inf_ms = timed(g.replay)
        assert step.overflowed_steps() == 0 and bool(torch.isfinite(logits).all())
        nvs = step.last_vertex_counts()
        execution = f"one CUDA graph per pass (static-shape lattice, rows per level {bounds})"
    else:
        fb_ms = timed(train_pass)
        loss = float(train_pass().item())
        inf_ms = timed(infer_pass)
        nvs = [int(l.nr_lattice_vertices()) for l in getattr(model, "last_level_lattices", [])]
    return {
        "scene": name, "impl": impl, "n_points": spec["n"], "sigma": spec["sigma"], "hash_table_capacity": spec["capacity"],
        "nr_classes": spec["nr_classes"], "vertices_per_level": nvs,
        "fwd_bwd_ms": fb_ms, "scans_per_s": 1e3 / fb_ms, "points_per_s": spec["n"] * 1e3 / fb_ms,
        "inference_ms": inf_ms, "inference_points_per_s": spec["n"] * 1e3 / inf_ms,
        "steps": steps, "warmup": warmup, "loss": loss, "data": "synthetic", "dtype": "f32",
        "execution": execution + ", L2 flushed before every timed pass, median of the passes",
        "conv": ("reference Python modules + reference kernels: im2row buffer + fp32 mm" if impl == "reference" else
                 {0: "fp32 FMA on CUDA cores", 1: "tcgen05 3xTF32 split, fp32 accumulate", 2: "tcgen05 TF32"}[precision]),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="both", choices=["kitti", "scannet", "both"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--conv-precision", type=int, default=1, choices=[0, 1, 2])
    ap.add_argument("--mode", default="graph", choices=["graph", "eager"])
    args = ap.parse_args()
    for name in (["kitti", "scannet"] if args.scene == "both" else [args.scene]):
        try:
            line = run_scene(name, args.impl, args.steps, args.warmup, args.conv_precision, args.mode)
        except Exception as exc:       # keep going: the other scene is still worth its line
            import traceback
            traceback.print_exc(file=sys.stderr)
            line = {"scene": name, "impl": args.impl, "error": f"{type(exc).__name__}: {exc}"}
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
