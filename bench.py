#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

metric   : scans/sec, forward + backward + optimizer step (LatticeNet training step)
workload : configs[1] -- LatticeNet ShapeNet-part segmentation (lnn_train_shapenet.cfg architecture:
           pointnet [16,32,64]->32, 3 levels, blocks [3,3,3]/1/[2,2,2]; 7 classes), synthetic clouds of
           2,048 points, sigma 0.05, hash capacity 60,000, loss 0.5*Lovasz + 0.5*NLL, AdamW(amsgrad).
           One scene per step per GPU; scene-parallel across GPUs with one flat NCCL all-reduce of the
           weight gradients per step ("weak" scaling: per-GPU work is fixed).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU).  Rank 0 prints ONE JSON line.
`value`  = scans/s with the clouds already resident in HBM.
`e2e`    = the same step driven from pinned HOST buffers: H2D of positions/values/labels and a D2H
           read of the loss inside the timed region, every step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NR_POINTS = 2048
NR_CLASSES = 7
SIGMA = 0.05
CAPACITY = 60000
POOL = 16            # distinct clouds cycled through (every step sees a different lattice)
L2_FLUSH_BYTES = 256 << 20


# ------------------------------------------------------------------------------------------------
def synthetic_cloud(seed):
    """ShapeNet-object-like cloud: points on the faces of a 0.8 x 0.3 x 0.4 box, random rigid jitter
    (translation +-0.2 in x,z) and 1 mm noise (config/lnn_train_shapenet.cfg:75-91)."""
    rng = np.random.RandomState(seed)
    size = np.array([0.8, 0.3, 0.4])
    p = (rng.rand(NR_POINTS, 3) - 0.5) * size
    face = rng.randint(0, 3, NR_POINTS)
    side = rng.randint(0, 2, NR_POINTS) * 2 - 1
    p[np.arange(NR_POINTS), face] = 0.5 * size[face] * side
    p += np.array([rng.uniform(-0.2, 0.2), 0.0, rng.uniform(-0.2, 0.2)])
    p += rng.randn(NR_POINTS, 3) * 0.001
    labels = rng.randint(0, NR_CLASSES, NR_POINTS)
    return p.astype(np.float32), labels.astype(np.int64)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu_index), "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, sm_max, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                sm_max = float(parts[2])
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower() == "active":
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": sm_max, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def build_training(device):
    from lattice_net_b200 import Lattice, ModelParams
    from lattice_net_b200.models import LNN
    lattice = Lattice(CAPACITY, [(SIGMA, 3)], name="lattice")
    model = LNN(NR_CLASSES, ModelParams(), device=device).to(device)
    return lattice, model


def train_step(model, lattice, pos, vals, labels, optimizer, bucket, world):
    from lattice_net_b200.losses import segmentation_loss
    logsoftmax, _ = model(lattice, pos, vals)
    loss = segmentation_loss(logsoftmax, labels)
    bucket.zero()
    loss.backward()
    bucket.allreduce_mean(world)
    optimizer.step()
    return loss


def timed_region(fn, steps, world, device, flush_buf):
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i in range(steps):
        flush_buf.fill_(float(i))          # evict L2 between steps (inside the timed region: ~40 us each)
        fn(i)
    end.record()
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    ms = torch.tensor([start.elapsed_time(end)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def kernel_roofline(device):
    """Live CUDA-event timing of the dominant lattice kernel of this workload: the level-1 lattice
    convolution (128 -> 128 channels, K = 9*128) of the decoder's ResnetBlocks -- the tcgen05 kernel ALONE, as it runs
    inside the step: filter slabs prepared beforehand (once per optimizer step), output taken from the zeroed arena."""
    import bench_ops
    from lattice_net_b200 import Lattice
    from lattice_net_b200 import lattice as lattice_mod
    pos = torch.from_numpy(synthetic_cloud(1234)[0]).to(device)
    lat = Lattice(CAPACITY, [(SIGMA, 3)])
    lat.begin_splat()
    lat.splat_standalone(pos, torch.zeros((NR_POINTS, 1), device=device))
    nv = lat.nr_lattice_vertices()
    cin = cout = 128
    F = 9
    lv = torch.randn((nv, cin), device=device)
    fb = torch.randn((F * cin, cout), device=device) * 0.05
    l2 = lat.clone_lattice()
    l2.set_values(lv)
    lattice_mod.prepare_filters([(fb, F, cin, cout, False)])
    arena = lattice_mod.ZeroArena(nv * cout + 64, device)
    prev_arena = lattice_mod.set_zero_arena(arena)

    def one():
        arena.off = 0          # the same (already accumulated-into) buffer again: timing only
        l2.convolve_im2row_standalone(fb, 1, l2, False)

    try:
        # device time only: the calls are captured into a CUDA graph (as in the benchmarked step) and the replay is timed
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(3):
                one()
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        per_graph = 20
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(per_graph):
                one()
    finally:
        lattice_mod.set_zero_arena(prev_arena)
    graph.replay()
    torch.cuda.synchronize(device)
    replays = 10
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(replays):
        graph.replay()
    ev[1].record()
    torch.cuda.synchronize(device)
    sec = ev[0].elapsed_time(ev[1]) * 1e-3 / (replays * per_graph)
    flops = 2.0 * nv * F * cin * cout
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops", 1590.0))
    tf32_peak = bench_ops.TF32 or bench_ops.measure_tf32_peak()
    achieved = flops / sec / 1e12
    # DRAM bytes of the same launch from one `ncu --set full` capture of exactly this configuration:
    # scripts/ncu_bench_conv.py -> scripts/roofline_traffic.py -> profiles/roofline_traffic.json
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            traffic = float(json.load(f)["traffic_bytes_per_call"])
    except Exception:
        pass
    return {"bound": "tensor", "kernel": "conv_tc3_kernel: lattice conv fwd 128->128 (K=1152), nv=%d, precision mode %d (3xTF32: three tcgen05 kind::tf32 passes per "
                                         "algorithmic flop); one launch = the tcgen05 kernel alone (filter slabs prepared once per step), device time from a CUDA-graph "
                                         "replay of 20 back-to-back launches (working set stays in L2, as inside the step)" % (nv, lattice_mod.CONV_PRECISION),
            "achieved": achieved, "peak": peak, "peak_source": "measured bf16 burst (MEASURED_PEAKS.json)" if peaks else "fallback",
            "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "us_per_launch": sec * 1e6,
            "tf32_peak_measured_here": tf32_peak, "frac_of_tf32_peak": achieved / tf32_peak,
            "algorithmic_flops": flops, "algorithmic_bytes": 4.0 * (nv * cin + nv * F + F * cin * cout + nv * cout)}


def extra_measurements(args):
    """BASELINE configs[2..4] next to the headline: SemanticKITTI- / ScanNet-sized scans (graph mode) and the operator
    roofline points at 10^6 points.  Rank 0, N = 1 only; ~1 minute."""
    import bench_ops
    import bench_scenes
    out = {}
    try:
        ops = bench_ops.headline_ops(1000000, (64, 128))
        out["ops"] = {"n_points": 1000000, "hbm_peak_GBps": bench_ops.HBM, "bf16_peak_TFLOPs": bench_ops.TF, "tf32_peak_TFLOPs_measured_here": bench_ops.TF32,
                      "entries": [{k: v for k, v in r.items() if k in ("op", "nv", "val_dim", "c_out", "us", "algo_GB", "GBps", "hbm_frac", "TFLOPs",
                                                                          "tensor_frac", "tf32_frac", "points_per_s")} for r in ops]}
    except Exception as exc:
        out["ops"] = {"error": f"{type(exc).__name__}: {exc}"}
    torch.cuda.empty_cache()
    scenes = []
    for name in ("kitti", "scannet"):
        try:
            r = bench_scenes.run_scene(name, "ours", 5, 2, args.conv_precision, "graph")
            scenes.append({k: r[k] for k in ("scene", "n_points", "vertices_per_level", "fwd_bwd_ms", "scans_per_s", "points_per_s", "inference_ms",
                                              "inference_points_per_s", "execution", "conv")})
        except Exception as exc:
            import traceback
            traceback.print_exc(file=sys.stderr)
            scenes.append({"scene": name, "error": f"{type(exc).__name__}: {exc}"})
        torch.cuda.empty_cache()
    out["scenes"] = scenes
    return out


def run_ours(args):
    from lattice_net_b200 import _cabi
    from lattice_net_b200.graphed import GraphedTrainStep, estimate_vertex_bounds
    from lattice_net_b200.losses import segmentation_loss
    from lattice_net_b200.parallel import GradBucket, broadcast_parameters, init_distributed
    rank, world, local_rank = init_distributed("nccl")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    torch.manual_seed(0)

    from lattice_net_b200 import set_conv_precision
    set_conv_precision(args.conv_precision)
    lattice, model = build_training(device)
    clouds = [synthetic_cloud(1000 * rank + i) for i in range(POOL)]
    dev_clouds = [(torch.from_numpy(p).to(device), torch.zeros((NR_POINTS, 1), device=device), torch.from_numpy(l).to(device)) for p, l in clouds]
    host_clouds = [(torch.from_numpy(p).pin_memory(), torch.zeros((NR_POINTS, 1)).pin_memory(), torch.from_numpy(l).pin_memory()) for p, l in clouds]

    # lazily created parameters exist after one forward; then lay out optimizer + gradient bucket
    with torch.no_grad():
        model(lattice, *dev_clouds[0][:2])
    broadcast_parameters(model, 0)
    graphed = args.mode == "graph"
    bucket = GradBucket(model.parameters())
    if graphed and not args.torch_optimizer:
        from lattice_net_b200.optim import FlatAdamW          # AdamW-amsgrad as one kernel over the flat parameter / gradient buffers
        optimizer = FlatAdamW(bucket, lr=1e-3, weight_decay=3e-4)
    else:
        optimizer = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=3e-4, amsgrad=True, fused=True, capturable=graphed)
    flush_buf = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device=device)

    bounds = None
    if graphed:
        # static-shape mode + one CUDA graph per step (lattice_net_b200/graphed.py); rows per lattice level from
        # the vertex counts of this rank's clouds (+30 %), identical on every replay
        nr_levels = model.nr_downsamples + 1
        bounds = estimate_vertex_bounds(CAPACITY, [(SIGMA, 3)], [c[0] for c in dev_clouds], nr_levels, headroom=1.3)
        step = GraphedTrainStep(model, lattice, optimizer, segmentation_loss, NR_POINTS, 3, 1, bounds, bucket, world,
                                warmup=3, capture_collective=args.capture_collective, example=dev_clouds[0],
                                overlap_allreduce=not args.no_overlap_allreduce)
        launches_per_step = step.launches_per_step

        def step_resident(i):
            step(*dev_clouds[i % POOL])

        losses = []

        from lattice_net_b200.data import PinnedCloudFeeder
        feeder = PinnedCloudFeeder(NR_POINTS, 3, 1, device)      # pinned double buffer: the H2D of cloud i+1 runs under step i
        feeder.stage(*host_clouds[0])

        def step_e2e(i):
            slot, (pos, vals, labels) = feeder.current()
            loss = step(pos, vals, labels)                           # asynchronous: input copies + one graph launch
            feeder.release(slot)
            feeder.stage(*host_clouds[(i + 1) % POOL])               # host -> pinned -> device copy of the NEXT cloud, under this step
            losses.append(float(loss.item()))                       # D2H read of the step's result
    else:
        def step_resident(i):
            pos, vals, labels = dev_clouds[i % POOL]
            train_step(model, lattice, pos, vals, labels, optimizer, bucket, world)

        losses = []

        def step_e2e(i):
            hp, hv, hl = host_clouds[i % POOL]
            pos = hp.to(device, non_blocking=True)
            vals = hv.to(device, non_blocking=True)
            labels = hl.to(device, non_blocking=True)
            loss = train_step(model, lattice, pos, vals, labels, optimizer, bucket, world)
            losses.append(float(loss.item()))      # D2H read of the step's result

    for i in range(args.warmup):
        step_resident(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    _cabi.reset_launch_count()
    ms = timed_region(step_resident, args.steps, world, device, flush_buf)
    launches = _cabi.launch_count() if not graphed else launches_per_step * args.steps
    # the headline region is K steps = a few tens of milliseconds: four more identical regions show how stable it is
    repeats = [ms] + [timed_region(step_resident, args.steps, world, device, flush_buf) for _ in range(4)]
    for i in range(max(1, args.warmup // 2)):
        step_e2e(i)
    ms_e2e = timed_region(step_e2e, args.steps, world, device, flush_buf)
    clocks = sampler.stop() if rank == 0 else None
    overflowed = step.overflowed_steps() if graphed else 0
    in_sync = None
    if world > 1:
        # every rank applied the same all-reduced gradients to the same broadcast parameters: the replicas must be bit-identical
        import torch.distributed as dist
        flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
        digest = torch.stack([flat.double().sum(), flat.double().abs().sum()])
        digests = [torch.empty_like(digest) for _ in range(world)]
        dist.all_gather(digests, digest)
        in_sync = all(bool(torch.equal(d, digests[0])) for d in digests)

    if rank != 0:
        return
    roof = kernel_roofline(device)
    cpu = cpu_baseline(model) if (world == 1 and not args.no_cpu_baseline) else None     # reported at N=1 only
    extras = extra_measurements(args) if (world == 1 and not args.no_extras) else {}
    scans = args.steps * world
    h2d = NR_POINTS * (3 * 4 + 1 * 4 + 8)
    line = {
        "metric": "scans/sec fwd+bwd", "value": scans / (ms * 1e-3), "unit": "scans/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "LatticeNet ShapeNet-part segmentation fwd+bwd+AdamW (lnn_train_shapenet.cfg arch), 2048-pt synthetic clouds, 1 scene/step/GPU",
                   "nr_points": NR_POINTS, "nr_classes": NR_CLASSES, "sigma": SIGMA, "hash_table_capacity": CAPACITY,
                   "parallelism": (f"scene-parallel dp{world}, {bucket.nbytes} gradient bytes all-reduced per step over NCCL"
                                   + (", in two chunks: decoder + slice head under the encoder's backward pass, the rest after it" if (graphed and getattr(step, "split", None) is not None) else ", one flat all-reduce")),
                   "replicas_bit_identical_after_run": in_sync,
                   "execution": ("%s (static-shape lattice, rows per level %s; %d step(s) skipped for exceeding them)"
                                 % ("one CUDA graph per step" + (", NCCL all-reduce captured inside it" if world > 1 else "") if len(step.graphs) == 1
                                    else "two CUDA graphs per step with an eager NCCL all-reduce between them", bounds, overflowed))
                                if graphed else "eager launches (dynamic-shape lattice)",
                   "l2": f"flushed between steps by a {L2_FLUSH_BYTES >> 20} MiB write (inside the timed region)",
                   "conv_precision": {0: "fp32 FMA on CUDA cores", 1: "tcgen05 3xTF32 split, fp32 accumulate (fp32-equivalent)", 2: "tcgen05 TF32"}[args.conv_precision]},
        "e2e": {"value": scans / (ms_e2e * 1e-3), "unit": "scans/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        "final_loss": losses[-1] if losses else None,
        "stability": {"ms_per_step_of_5_regions": [r / args.steps for r in repeats], "median_value": scans / (float(np.median(repeats)) * 1e-3),
                      "note": "`value` is the FIRST region (the contract's K timed steps); the other four follow it back to back"},
    }
    line.update(extras)
    print(json.dumps(line))


def cpu_baseline(model):
    try:
        from oracle import cpu_port
    except Exception as exc:     # the oracle is test infrastructure; say so instead of failing the bench
        return {"value": None, "unit": "scans/s", "cores": 0, "kind": "port", "sample": f"unavailable: {exc}"}
    return cpu_port.time_training_step(model, synthetic_cloud, NR_CLASSES, SIGMA, budget_s=20.0)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    try:
        from oracle import ref_arm
        line = ref_arm.run(args, synthetic_cloud, dict(nr_points=NR_POINTS, nr_classes=NR_CLASSES, sigma=SIGMA, capacity=CAPACITY))
    except Exception as exc:
        line = {"impl": "reference", "unavailable": f"{type(exc).__name__}: {exc}"}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="graph", choices=["graph", "eager"],
                    help="graph: static-shape lattice + one CUDA graph per step (default); eager: dynamic-shape launches")
    ap.add_argument("--capture-collective", dest="capture_collective", action="store_true", default=True,
                    help="N>1: capture the NCCL all-reduce inside the step graph (default; one graph launch per step)")
    ap.add_argument("--no-capture-collective", dest="capture_collective", action="store_false",
                    help="N>1: two graphs per step with an eager NCCL all-reduce between them")
    ap.add_argument("--conv-precision", type=int, default=1, choices=[0, 1, 2],
                    help="0 fp32 CUDA cores, 1 tcgen05 3xTF32 (fp32-equivalent, default), 2 tcgen05 TF32")
    ap.add_argument("--no-overlap-allreduce", action="store_true",
                    help="N>1: one all-reduce after the backward pass instead of two chunks, the late one under the encoder's backward")
    ap.add_argument("--no-extras", action="store_true", help="skip the `ops` / `scenes` keys (scene-sized scans and operator roofline points, ~1 min)")
    ap.add_argument("--torch-optimizer", action="store_true", help="torch.optim.AdamW(fused) instead of the one-kernel flat AdamW")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the ~20 s CPU leg (profiler passes only; never for a reported line)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    # no destroy_process_group(): with NCCL collectives captured inside live CUDA graphs it blocked at exit on 8 GPUs (r02w);
    # the process group is torn down with the interpreter, as in round 1


if __name__ == "__main__":
    main()
