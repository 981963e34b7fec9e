#!/usr/bin/env python
"""Operator micro-benchmarks (BASELINE.json configs[4]: the splat / slice / conv sweep).

For every op: CUDA-event time of our kernel, achieved algorithmic HBM GB/s (bytes of SURVEY.md
section 8d) or TFLOP/s, the fraction of the measured peak (MEASURED_PEAKS.json), and -- where the
reference's own kernel is available (oracle/_ref) -- the reference kernel's time on the same inputs.
Inputs are sized well above the 126 MB L2 or the L2 is flushed between iterations.

    python bench_ops.py [--n 1000000] [--quick] > ops.jsonl
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p["bf16_tflops"]), "measured"
    except Exception:
        return 6650.0, 1590.0, "fallback"


HBM, TF, PEAK_SRC = peaks()
TF32 = None        # dense TF32 peak measured on this GPU (measure_tf32_peak), the denominator of the tcgen05 kind::tf32 kernels
_flush = None
_SINK = None       # list: emit() collects records here instead of printing (bench.py's `ops` key)


def measure_tf32_peak(n=8192, reps=6):
    """Dense TF32 GEMM throughput of this GPU (cuBLAS through torch.matmul with TF32 allowed), best of `reps`:
    BASELINE.md section 2 asks for a measured TF32 denominator next to the bf16 one."""
    global TF32
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn((n, n), device="cuda")
        b = torch.randn((n, n), device="cuda")
        torch.matmul(a, b)
        torch.cuda.synchronize()
        best = float("inf")
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e-3)
        TF32 = 2.0 * n ** 3 / best / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    return TF32


def timeit(fn, reps=10, warm=3, flush=True):
    global _flush
    if _flush is None:
        _flush = torch.empty((256 << 20) // 4, dtype=torch.float32, device="cuda")
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for i in range(reps):
        if flush:
            _flush.fill_(float(i))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return float(np.median(ts))


def emit(op, cfg, sec, nbytes=None, flops=None, ref_sec=None, extra=None):
    rec = {"op": op, **cfg, "us": sec * 1e6}
    if nbytes is not None:
        rec.update(algo_GB=nbytes / 1e9, GBps=nbytes / sec / 1e9, hbm_frac=nbytes / sec / 1e9 / HBM)
    if flops is not None:
        rec.update(TFLOPs=flops / sec / 1e12, tensor_frac=flops / sec / 1e12 / TF)       # of the measured bf16 dense peak
        if TF32:
            rec.update(tf32_frac=flops / sec / 1e12 / TF32)                                # of the measured TF32 dense peak
    if ref_sec is not None:
        rec.update(ref_us=ref_sec * 1e6, speedup_vs_ref_kernel=ref_sec / sec)
    if extra:
        rec.update(extra)
    rec["peak_source"] = PEAK_SRC
    if _SINK is not None:
        _SINK.append(rec)
    else:
        print(json.dumps(rec), flush=True)


def uniform_cloud(n, d, seed, order="random"):
    """n points uniform in the unit cube (sigma is then tuned so that nv ~ N/2).
    order="morton": the same points sorted along a Z-order curve of their first three coordinates -- the
    ordering a spatially coherent scan (lidar sweep, mesh sampling) has; neighbouring points then share
    lattice vertices, which is what lets the gathers hit in L1 instead of L2."""
    rng = np.random.RandomState(seed)
    pos = rng.rand(n, d).astype(np.float32)
    if order == "morton":
        q = np.minimum((pos[:, :3] * 1024).astype(np.uint64), 1023)
        code = np.zeros(n, np.uint64)
        for bit in range(10):
            for a in range(3):
                code |= ((q[:, a] >> np.uint64(bit)) & np.uint64(1)) << np.uint64(3 * bit + a)
        pos = pos[np.argsort(code, kind="stable")]
    return np.ascontiguousarray(pos)


def bench_group_norm(nv, widths, cfg0, dev):
    """GroupNorm + ReLU between the convolutions: this repo's kernels vs what the reference runs
    (transpose to [1, C, nv], torch.nn.GroupNorm, ReLU, transpose back: lattice_modules.py:585-614)."""
    from lattice_net_b200.lattice_modules import _GroupNormReLU
    for C in widths:
        xg = torch.randn((nv, C), device=dev)
        gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        gy = torch.randn((nv, C), device=dev)
        cfg = dict(cfg0, val_dim=C)

        def gn_ref():
            return torch.relu(torch.nn.functional.group_norm(xg.t().unsqueeze(0), 32, gamma, beta, 1e-5)).squeeze(0).t().contiguous()
        emit("group_norm_relu_fwd", cfg, timeit(lambda: _GroupNormReLU.apply(xg, gamma, beta, 32, 1e-5, True), reps=5),
             nbytes=8.0 * nv * C, ref_sec=timeit(gn_ref, reps=5), extra={"ref": "torch GroupNorm + ReLU + layout copies"})
        xr = xg.clone().requires_grad_(True)
        yr = _GroupNormReLU.apply(xr, gamma, beta, 32, 1e-5, True)
        emit("group_norm_relu_bwd", cfg, timeit(lambda: torch.autograd.grad(yr, xr, gy, retain_graph=True), reps=5), nbytes=16.0 * nv * C)
        del xg, gy, xr, yr


def run(n, d, vals, quick, order="random", with_ref=True, with_extras=True):
    from lattice_net_b200 import Lattice, lattice as lm
    from lattice_net_b200._cabi import call, ptr, stream_ptr
    from oracle import ref_cuda
    have_ref = with_ref and ref_cuda.available()
    dev = torch.device("cuda", 0)
    # sigma so that nv ~ n/2: vertices ~ (d+1)! * volume/sigma^d density; tune by a coarse search
    pos = torch.from_numpy(uniform_cloud(n, d, 0, order)).to(dev)
    cap = int(4 * n)
    sigma = (1.0 / n) ** (1.0 / d) * (2.2 if d == 3 else 1.6)
    for _ in range(6):
        lat = Lattice(cap, [(sigma, d)])
        lat.begin_splat()
        idx, w = lat.just_create_verts(pos, True)
        nv = lat.nr_lattice_vertices()
        if nv > 0.65 * n:
            sigma *= 1.25
        elif nv < 0.35 * n:
            sigma *= 0.85
        else:
            break
    cfg0 = {"n": n, "pos_dim": d, "nv": nv, "capacity": cap, "sigma": round(sigma, 5), "point_order": order}
    st = lat.hash_table().structure
    sig = lat._sigmas_on(dev)

    def build():
        st.clear()
        call("ln_splat_build", ptr(pos), ptr(sig), n, d, ptr(st.keys), ptr(st.entries), ptr(st.nr_filled), ptr(st.status), st.capacity, 0, ptr(idx), ptr(w), stream_ptr(dev))

    ref = None
    ref_sec = None
    if have_ref:
        ref = ref_cuda.RefLattice(cap, [sigma] * d)
        ref.table = ref_cuda.RefTable(cap, d)

        def ref_build():
            ref.table.entries.fill_(-1)
            ref.table.nr_filled.fill_(0)
            return ref.build(pos)
        ref_sec = timeit(ref_build, reps=5)
        ridx, rw = ref_build()
    sec = timeit(build, reps=5)
    st.mark_dirty()
    nv = lat.nr_lattice_vertices()
    emit("splat_build(+table clear)", cfg0, sec, nbytes=4 * n * d + 8 * n * (d + 1) + 4 * nv * d, ref_sec=ref_sec,
         extra={"points_per_s": n / sec, "max_probe": st.max_probe})

    for V in vals:
        cfg = dict(cfg0, val_dim=V)
        x = torch.randn((n, V), device=dev)
        lv = torch.zeros((nv, V), device=dev)

        def acc():
            call("ln_splat_accumulate", ptr(x), ptr(idx), ptr(w), n, d, V, nv, ptr(lv), stream_ptr(dev))
        rs = None
        if ref is not None and ref.k.has(f"splatCacheNaive<{d},{V}>"):
            rvals = torch.zeros((cap, V), device=dev)
            import ctypes
            rs = timeit(lambda: ref.k.launch(f"splatCacheNaive<{d},{V}>", n, [ctypes.c_int(n), ref_cuda._p(x), ref_cuda._p(ridx), ref_cuda._p(rw), ref.table.struct(rvals)]), reps=5)
            del rvals
        emit("splat_accumulate", cfg, timeit(acc, reps=5), nbytes=4 * n * V + 8 * n * (d + 1) + 4 * nv * V, ref_sec=rs)

        lat2 = lat.clone_lattice()
        lvr = torch.randn((nv, V), device=dev)
        lat2.set_values(lvr)
        out = torch.empty((n, V), device=dev)

        def sl():
            call("ln_slice_fwd", ptr(lvr), ptr(idx), ptr(w), n, d, V, nv, ptr(out), stream_ptr(dev))
        rs = None
        if ref is not None and ref.k.has(f"slice_with_precomputation<{d},{V}>"):
            rs = timeit(lambda: ref.slice_with_precomputation(pos, lvr, ridx, rw), reps=5)
        emit("slice_fwd", cfg, timeit(sl, reps=5), nbytes=8 * n * (d + 1) + 4 * nv * V + 4 * n * V, ref_sec=rs,
             extra={"gather_GB": 4 * n * (d + 1) * V / 1e9})
        g = torch.randn((n, V), device=dev)
        gl = torch.zeros((nv, V), device=dev)

        def slb():
            call("ln_slice_bwd", ptr(g), ptr(idx), ptr(w), n, d, V, nv, ptr(gl), stream_ptr(dev))
        rs = None
        if ref is not None and ref.k.has(f"slice_backwards_with_precomputation_no_homogeneous<{d},{V}>"):
            rs = timeit(lambda: ref.slice_backwards(g, ridx, rw), reps=5)
        emit("slice_bwd", cfg, timeit(slb, reps=5), nbytes=8 * n * (d + 1) + 4 * nv * V + 4 * n * V, ref_sec=rs)
        del x, lv, out, g, gl

        if V >= 32:
            F = 2 * (d + 1) + 1
            fb = torch.randn((F * V, V), device=dev) * 0.05
            lat2._neighbour_table(lat2, 1)
            flops = 2.0 * nv * F * V * V
            cbytes = 4.0 * (nv * V + nv * F + F * V * V + nv * V)
            for prec, name in ((0, "conv_fwd fp32 SIMT"), (1, "conv_fwd tcgen05 3xTF32"), (2, "conv_fwd tcgen05 TF32")):
                if prec == 0 and quick:
                    continue
                lm.set_conv_precision(prec)
                try:
                    lm.prepare_filters([(fb, F, V, V, False)])       # filter slabs are prepared once per optimizer step, not per call
                    s = timeit(lambda: lat2.convolve_im2row_standalone(fb, 1, lat2, False), reps=5)
                    emit(name, dict(cfg, c_out=V), s, nbytes=cbytes, flops=flops)
                finally:
                    lm.set_conv_precision(1)
            if ref is not None and ref.k.has(f"im2row<{d},{V}>"):
                rs = timeit(lambda: ref.convolve(fb, ref, lvr, 1, False), reps=3)
                emit("conv_fwd reference (im2row kernel + fp32 mm)", dict(cfg, c_out=V), rs, nbytes=cbytes, flops=flops)
            gout = torch.randn((nv, V), device=dev)
            for prec, name in ((0, "conv_wgrad fp32 SIMT"), (1, "conv_wgrad tcgen05 3xTF32"), (2, "conv_wgrad tcgen05 TF32")):
                if prec == 0 and quick:
                    continue
                lm.set_conv_precision(prec)
                try:
                    emit(name, dict(cfg, c_out=V), timeit(lambda: lat2.conv_weight_grad(lat2, gout, F, 1), reps=5), nbytes=cbytes, flops=flops)
                finally:
                    lm.set_conv_precision(1)
            del fb, gout
        del lat2, lvr
    if not with_extras:
        return
    bench_group_norm(nv, [v for v in vals if v >= 32][:2], cfg0, dev)
    # fused slice + classify, widths for which the reference instantiation is built
    for V, nc in [(v, c) for v, c in ((32, 7), (64, 16), (128, 20)) if v in vals]:
        cfg = dict(cfg0, val_dim=V, nr_classes=nc)
        lvc = torch.randn((nv, V), device=dev)
        dw = torch.randn((n, d + 1), device=dev) * 0.05
        cw, cb = torch.randn((nc, V), device=dev) * 0.2, torch.zeros((nc,), device=dev)
        gl = torch.randn((n, nc), device=dev)
        latc = lat.clone_lattice()
        latc.set_values(lvc)
        nbytes = 12.0 * n * (d + 1) + 4.0 * nv * V + 4.0 * n * nc + 4.0 * nc * V
        rs_f = rs_b = None
        if ref is not None and ref.k.has(f"slice_classify_with_precomputation<{d},{V},{nc}>"):
            try:
                lvr_c = torch.randn((ref.nv(), V), device=dev)
                rs_f = timeit(lambda: ref.slice_classify_with_precomputation(pos, lvr_c, dw, cw, cb, ridx, rw), reps=3)
                rs_b = timeit(lambda: ref.slice_classify_backwards(gl, lvr_c, dw, cw, cb, ridx, rw), reps=3)
            except Exception as exc:      # the sweep goes on without the reference column
                print(f"# reference slice_classify<{d},{V},{nc}> not timed: {exc}", file=sys.stderr)
        emit("slice_classify_fwd", cfg, timeit(lambda: latc.slice_classify_with_precomputation(pos, dw, cw, cb, nc, idx, w), reps=5),
             nbytes=nbytes, ref_sec=rs_f)
        g_lv, g_dw, g_w, g_b = torch.zeros_like(lvc), torch.zeros_like(dw), torch.zeros_like(cw), torch.zeros_like(cb)
        emit("slice_classify_bwd", cfg, timeit(lambda: latc.slice_classify_backwards_with_precomputation(gl, pos, lvc, dw, cw, cb, nc, g_lv, g_dw, g_w, g_b, idx, w), reps=5),
             nbytes=nbytes + 4.0 * nv * V + 4.0 * n * (d + 1), ref_sec=rs_b)
        del lvc, dw, gl, g_lv, g_dw, latc
    # neighbour table
    lat3 = lat.clone_lattice()
    lat3.set_values(torch.zeros((nv, 1), device=dev))
    F = 2 * (d + 1) + 1

    def nt():
        st.neighbour_cache.clear()
        lat3._neighbour_table(lat3, 1)
    emit("neighbour_table", cfg0, timeit(nt, reps=5), nbytes=4 * nv * d + 4 * nv * F + 4 * nv * (d + 1))


def headline_ops(n=1000000, vals=(64, 128)):
    """The `ops` key of bench.py's JSON line: splat / slice / scatter / conv forward / weight gradient at n points, each with
    its algorithmic GB/s or TFLOP/s and the fraction of the measured peak."""
    global _SINK
    _SINK = []
    try:
        measure_tf32_peak()
        run(n, 3, list(vals), True, with_ref=False, with_extras=False)
        return _SINK
    finally:
        _SINK = None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, nargs="*", default=[100000, 1000000])
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--vals", type=int, nargs="*", default=None)
    ap.add_argument("--order", default="random", choices=["random", "morton"], help="point order of the synthetic cloud")
    ap.add_argument("--gn-only", type=int, default=0, metavar="NV", help="only the GroupNorm entries, on an [NV x C] tensor")
    ap.add_argument("--no-ref", action="store_true", help="skip the reference-kernel columns (their im2row buffer is F x the activations)")
    ap.add_argument("--pos-dims", type=int, nargs="*", default=None, help="position dimensions to sweep (default: 3, then 5 at half the last n)")
    args = ap.parse_args()
    measure_tf32_peak()
    print(json.dumps({"op": "peaks", "hbm_GBps": HBM, "bf16_TFLOPs": TF, "tf32_TFLOPs_measured_here": TF32, "peak_source": PEAK_SRC}), flush=True)
    if args.gn_only:
        bench_group_norm(args.gn_only, args.vals or [32, 64], {"n": 0, "pos_dim": 3, "nv": args.gn_only, "point_order": "-"}, torch.device("cuda", 0))
        return
    if args.pos_dims:
        for d in args.pos_dims:
            for n in args.n:
                run(n, d, args.vals if args.vals else [8, 32, 64, 128, 256], args.quick, args.order, with_ref=not args.no_ref and n <= 1000000)
        return
    for n in args.n:
        run(n, 3, args.vals if args.vals else ([8, 32, 64] if args.quick else [1, 8, 32, 64, 128]), args.quick, args.order)
    if not args.quick:
        run(args.n[-1] // 2, 5, [8, 32], True)


if __name__ == "__main__":
    main()
