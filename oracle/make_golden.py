#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by running the REFERENCE's own kernels
(oracle/_ref, see oracle/build_ref.py) on a B200 with the seeded inputs of oracle/cases.py.

Run on the GPU box:   gpurun -- python oracle/make_golden.py gpurun_out/golden
then copy gpurun_out/golden/* into tests/golden/ and commit.  Vertex ids are race-dependent in the
reference (atomicAdd order, HashTableGPU.cuh:454); everything is stored in the canonical numbering
(keys sorted lexicographically) so fixtures are reproducible.
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases, lattice_oracle as lo   # noqa: E402
from oracle.ref_cuda import RefKernels, RefLattice, REF_DIR   # noqa: E402


def probe_rsqrt(out_dir):
    from cuda.bindings import driver
    torch.zeros(1, device="cuda")
    with open(os.path.join(REF_DIR, "probe.ptx"), "rb") as f:
        err, mod = driver.cuModuleLoadData(f.read() + b"\0")
    assert int(err) == 0
    err, fn = driver.cuModuleGetFunction(mod, b"probe_rsqrt")
    assert int(err) == 0
    xs = torch.tensor([2.0, 6.0, 12.0, 20.0, 30.0], device="cuda")
    ys = torch.zeros_like(xs)
    args = [ctypes.c_void_p(xs.data_ptr()), ctypes.c_void_p(ys.data_ptr()), ctypes.c_int(5)]
    ptrs = (ctypes.c_void_p * 3)(*[ctypes.addressof(a) for a in args])
    (err,) = driver.cuLaunchKernel(fn, 1, 1, 1, 32, 1, 1, 0, torch.cuda.current_stream().cuda_stream, ctypes.addressof(ptrs), 0)
    assert int(err) == 0
    torch.cuda.synchronize()
    bits = [f"0x{int(b):08X}" for b in ys.cpu().numpy().view(np.uint32)]
    with open(os.path.join(out_dir, "rsqrt_approx.json"), "w") as f:
        json.dump({"args": [2, 6, 12, 20, 30], "bits": bits, "gpu": torch.cuda.get_device_name(0),
                   "what": "rsqrt.approx.ftz.f32 results used by the reference's elevate()"}, f, indent=1)
    print("rsqrt.approx.ftz bits:", bits)


def canon(keys_gpu, nv):
    keys = keys_gpu[:nv].cpu().numpy()
    return lo.canonical_order(keys)


def gen_case(name, spec, out_dir):
    pos_np = spec["make"]()
    n, d = pos_np.shape
    pos = torch.from_numpy(pos_np).cuda()
    out = dict(positions=pos_np, sigmas=np.asarray(spec["sigmas"], np.float32), capacity=np.int32(spec["capacity"]))
    F = 2 * (d + 1) + 1

    # ---- structure + splat (kernel_splat, splatCacheNaive) -------------------------------------------
    V = 3
    vals_np = cases.randn((n, V), 10)
    lat = RefLattice(spec["capacity"], spec["sigmas"])
    idx, w = lat.splat(pos, torch.from_numpy(vals_np).cuda())
    nv = lat.nv()
    keys_sorted, o2n = canon(lat.table.keys, nv)
    n2o = np.argsort(o2n)
    out.update(nv=np.int32(nv), keys=keys_sorted.astype(np.int32), indices=lo.relabel(idx.cpu().numpy(), o2n).astype(np.int32),
               weights=w.cpu().numpy(), splat_in=vals_np, splat_values=lat.values[:nv].cpu().numpy()[n2o])
    print(f"[{name}] n={n} d={d} nv={nv}")

    # ---- distribute -----------------------------------------------------------------------------------
    dl = RefLattice(spec["capacity"], spec["sigmas"])
    dv_np = cases.randn((n, 1), 11)
    distributed, didx, dw = dl.distribute(pos, torch.from_numpy(dv_np).cuda())
    dk, do2n = canon(dl.table.keys, dl.nv())
    assert np.array_equal(dk, keys_sorted)
    out.update(distribute_in=dv_np, distributed=distributed.cpu().numpy(),
               distribute_indices=lo.relabel(didx.cpu().numpy(), do2n).astype(np.int32), distribute_weights=dw.cpu().numpy())

    # lattice values defined in canonical order; the GPU table wants them in its own order
    def to_gpu_order(lv_canon, o2n_):
        return torch.from_numpy(lv_canon[o2n_]).cuda().contiguous()

    # ---- slice / gather / slice_classify + backwards ------------------------------------------------------
    for Vs in (1, 8, 32):
        lv = cases.randn((nv, Vs), 20 + Vs)
        lvg = to_gpu_order(lv, o2n)
        out[f"lv{Vs}"] = lv
        out[f"slice{Vs}"] = lat.slice_with_precomputation(pos, lvg, idx, w).cpu().numpy()
        g = cases.randn((n, Vs), 30 + Vs)
        out[f"slice_bwd_in{Vs}"] = g
        out[f"slice_bwd{Vs}"] = lat.slice_backwards(torch.from_numpy(g).cuda(), idx, w).cpu().numpy()[n2o]
    s_np, i_np, w_np = [t.cpu().numpy() for t in lat.slice_no_precomputation(pos * 1.0, to_gpu_order(out["lv8"], o2n))]
    out.update(slice_nop8=s_np, slice_nop_indices=lo.relabel(i_np, o2n).astype(np.int32), slice_nop_weights=w_np)
    # positions slightly off the cloud: some simplex vertices do not exist -> -1 entries
    pos_off = torch.from_numpy((pos_np + np.float32(0.013)).astype(np.float32)).cuda()
    s_np, i_np, w_np = [t.cpu().numpy() for t in lat.slice_no_precomputation(pos_off, to_gpu_order(out["lv8"], o2n))]
    out.update(slice_off8=s_np, slice_off_indices=lo.relabel(i_np, o2n).astype(np.int32), slice_off_weights=w_np)

    lv8g = to_gpu_order(out["lv8"], o2n)
    out["gather8"] = lat.gather_with_precomputation(pos, lv8g, idx, w).cpu().numpy()
    gg = cases.randn((n, (d + 1) * 9), 40)
    out["gather_bwd_in8"] = gg
    out["gather_bwd8"] = lat.gather_backwards(torch.from_numpy(gg).cuda(), idx, w).cpu().numpy()[n2o]

    if d == 3:
        Vc, nc = 32, 7
        lvc = out["lv32"]
        dw_np = (cases.randn((n, d + 1), 50) * 0.05).astype(np.float32)
        cw = (cases.randn((nc, Vc), 51) * 0.2).astype(np.float32)
        cb = cases.randn((nc,), 52)
        gl = cases.randn((n, nc), 53)
        tg = lambda a: torch.from_numpy(a).cuda()
        logits = lat.slice_classify_with_precomputation(pos, to_gpu_order(lvc, o2n), tg(dw_np), tg(cw), tg(cb), idx, w)
        g_lv, g_dw, g_w, g_b = lat.slice_classify_backwards(tg(gl), to_gpu_order(lvc, o2n), tg(dw_np), tg(cw), tg(cb), idx, w)
        out.update(sc_dw=dw_np, sc_w=cw, sc_b=cb, sc_grad_in=gl, sc_logits=logits.cpu().numpy(),
                   sc_g_lv=g_lv.cpu().numpy()[n2o], sc_g_dw=g_dw.cpu().numpy(), sc_g_w=g_w.cpu().numpy(), sc_g_b=g_b.cpu().numpy())

    # ---- neighbourhood: same level (dilation 1, 2), coarse levels --------------------------------------------
    def canon_rowindices(rowified, o2n_nbr, q_n2o, val_dim=1):
        # raw 0 is ambiguous in the reference (vertex id 0, or a cell its kernel never wrote into the
        # zeros-initialised buffer, Lattice.cu:600): store it as -3 and record which canonical vertex is id 0
        r = rowified.cpu().numpy().reshape(-1, F, val_dim)[:, :, 0]
        out_ = lo.relabel(r, o2n_nbr).astype(np.int32)
        out_[r == 0] = -3
        return out_[q_n2o]

    for dil in (1, 2):
        for flip in (False, True):
            ri = lat.im2rowindices(lat, 1, dil, flip)
            out[f"rowidx_d{dil}_f{int(flip)}"] = canon_rowindices(ri, o2n, n2o)
    Cin, Cout = 8, 16
    fb = (cases.randn((F * Cin, Cout), 60) * 0.1).astype(np.float32)
    out["conv_filter"] = fb
    conv = lat.convolve(torch.from_numpy(fb).cuda(), lat, lv8g, 1, False)
    out["conv8_16"] = conv.cpu().numpy()[n2o]
    rows = lat.im2row(lat, lv8g, 1, False)
    out["row2im8"] = lat.row2im(rows, lat, 8, 1).cpu().numpy()[n2o]

    coarse = lat.create_coarse_verts_naive(pos)
    nvc = coarse.nv()
    ck, co2n = canon(coarse.table.keys, nvc)
    cn2o = np.argsort(co2n)
    out.update(coarse_nv=np.int32(nvc), coarse_keys=ck.astype(np.int32), id0_fine=np.int64(o2n[0]), id0_coarse=np.int64(co2n[0]))
    # coarse query <- fine neighbours (coarsen fwd), fine query <- coarse neighbours (finefy fwd / coarsen bwd)
    out["rowidx_coarse_from_fine"] = canon_rowindices(coarse.im2rowindices(lat, 1, 1, False), o2n, cn2o)
    out["rowidx_fine_from_coarse"] = canon_rowindices(lat.im2rowindices(coarse, 1, 1, False), co2n, n2o)
    out["rowidx_fine_from_coarse_flip"] = canon_rowindices(lat.im2rowindices(coarse, 1, 1, True), co2n, n2o)
    conv_c = coarse.convolve(torch.from_numpy(fb).cuda(), lat, lv8g, 1, False)
    out["coarsen_conv8_16"] = conv_c.cpu().numpy()[cn2o]

    kc = lat.create_coarse_verts()
    kk, _ = canon(kc.table.keys, kc.nv())
    out["coarsen_kernel_keys"] = kk.astype(np.int32)

    np.savez_compressed(os.path.join(out_dir, f"{name}.npz"), **out)
    print(f"[{name}] wrote {len(out)} arrays, coarse nv={nvc}, coarsen-kernel nv={kc.nv()}")


def main():
    out_dir = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden"
    os.makedirs(out_dir, exist_ok=True)
    RefKernels.get()
    probe_rsqrt(out_dir)
    for name, spec in cases.CASES.items():
        gen_case(name, spec, out_dir)


if __name__ == "__main__":
    main()
