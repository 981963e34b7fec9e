"""Parity tests proper: the CUDA path, called through the C ABI via the `Lattice` handle, against
  (1) the reference's own kernels run live on the same GPU (oracle/ref_cuda.py), and
  (2) the CPU oracle (oracle/lattice_oracle.py) / the committed golden vectors.
Integers (keys, vertex counts, index tables, neighbour tables) must be bit-exact after the canonical
key sort; weights are expected bit-equal; accumulated values within the stated fp32 tolerances.
"""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import cases, lattice_oracle as lo
from tests.util import agg, assert_close, bits_equal, canonical, first, max_rel_err

pytestmark = pytest.mark.gpu

TOL_VALUES = 1e-4   # north_star: max rel-err <= 1e-4 on values (atomic / accumulation order differs)
TOL_GRADS = 1e-3    # and <= 1e-3 on gradients


def _lattice(spec):
    from lattice_net_b200 import Lattice
    d = len(spec["sigmas"])
    return Lattice(spec["capacity"], [(s, 1) for s in spec["sigmas"]])


def _ref():
    from oracle import ref_cuda
    if not ref_cuda.available():
        pytest.skip("oracle/_ref not built")
    return ref_cuda


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


_DEFAULT_PRECISION = 1      # the product default (lattice.py: CONV_PRECISION): tcgen05 3xTF32


def test_default_convolution_runs_on_the_tensor_cores():
    from lattice_net_b200 import lattice as lattice_mod
    assert lattice_mod.CONV_PRECISION == _DEFAULT_PRECISION == 1


@pytest.fixture(scope="module", params=list(cases.CASES))
def built(request):
    """Build the same lattice with our kernels, the reference kernels and the CPU oracle."""
    name = request.param
    spec = cases.CASES[name]
    pos_np = spec["make"]()
    pos = cuda(pos_np)
    n, d = pos_np.shape
    vals_np = cases.randn((n, 3), 10)
    ours = _lattice(spec)
    ours.begin_splat()
    idx, w = ours.splat_standalone(pos, cuda(vals_np))
    nv = ours.nr_lattice_vertices()
    keys = ours.hash_table().m_keys_tensor[:nv].cpu().numpy()
    ks, o2n, n2o = canonical(keys)
    ref_mod = _ref()
    ref = ref_mod.RefLattice(spec["capacity"], spec["sigmas"])
    ridx, rw = ref.splat(pos, cuda(vals_np))
    rnv = ref.nv()
    rks, ro2n, rn2o = canonical(ref.table.keys[:rnv].cpu().numpy())
    cpu = lo.build_lattice(pos_np, spec["sigmas"])
    return dict(name=name, spec=spec, pos_np=pos_np, pos=pos, n=n, d=d, vals_np=vals_np, ours=ours, idx=idx, w=w, nv=nv,
                ks=ks, o2n=o2n, n2o=n2o, ref=ref, ridx=ridx, rw=rw, rnv=rnv, rks=rks, ro2n=ro2n, rn2o=rn2o, cpu=cpu)


def test_structure_bit_exact(built):
    b = built
    assert b["nv"] == b["cpu"]["nv"] == len(b["rks"])
    if b["rnv"] != len(b["rks"]):
        print(f"note: the reference kernel stored {b['rnv'] - len(b['rks'])} duplicate vertices in this run")
    assert np.array_equal(b["ks"], b["rks"]), "key set differs from the reference kernels"
    assert np.array_equal(b["ks"], b["cpu"]["keys"]), "key set differs from the CPU oracle"
    ours_idx = lo.relabel(b["idx"].cpu().numpy(), b["o2n"])
    ref_idx = lo.relabel(b["ridx"].cpu().numpy(), b["ro2n"])
    assert np.array_equal(ours_idx, ref_idx), f"{(ours_idx != ref_idx).sum()} splatting indices differ from the reference"
    assert np.array_equal(ours_idx, b["cpu"]["indices"])
    assert bits_equal(b["w"].cpu().numpy(), b["rw"].cpu().numpy()) == 0, "barycentric weights are not bit-equal to the reference"
    assert bits_equal(b["w"].cpu().numpy(), b["cpu"]["weights"]) == 0, "barycentric weights are not bit-equal to the CPU oracle"
    # compact ids: a permutation of 0..nv-1, every key exactly once
    assert len(np.unique(b["ks"], axis=0)) == b["nv"]


def test_splat_values(built):
    b = built
    ours = b["ours"].values()
    assert tuple(ours.shape) == (b["spec"]["capacity"], 3)           # [capacity x V], like the reference
    ours_np = ours[:b["nv"]].cpu().numpy()[b["n2o"]]
    ref_np = agg(b["ref"].values[:b["rnv"]].cpu().numpy(), b["ro2n"])
    assert_close(ours_np, ref_np, TOL_VALUES, "splat values vs reference kernels")
    assert_close(ours_np, lo.splat_accumulate(b["vals_np"], b["cpu"]["indices"], b["cpu"]["weights"], b["nv"]), TOL_VALUES, "splat values vs oracle")
    assert float(ours[b["nv"]:].abs().max()) == 0.0


@pytest.mark.parametrize("V", [1, 4, 8, 32, 64, 3])
def test_splat_accumulate_widths(built, V):
    b = built
    from lattice_net_b200._cabi import call, ptr, stream_ptr
    vals = cases.randn((b["n"], V), 100 + V)
    out = torch.zeros((b["nv"], V), device="cuda")
    call("ln_splat_accumulate", ptr(cuda(vals)), ptr(b["idx"]), ptr(b["w"]), b["n"], b["d"], V, b["nv"], ptr(out), stream_ptr())
    exp = lo.splat_accumulate(vals, b["idx"].cpu().numpy(), b["w"].cpu().numpy(), b["nv"])
    assert_close(out.cpu().numpy(), exp, TOL_VALUES, f"splat accumulate V={V}")


def test_distribute(built):
    b = built
    spec, pos = b["spec"], b["pos"]
    dv = cases.randn((b["n"], 1), 11)
    ours = _lattice(spec)
    ours.begin_splat()
    dl, distributed, idx, w = ours.distribute(pos, cuda(dv))
    ref = _ref().RefLattice(spec["capacity"], spec["sigmas"])
    rdist, ridx, rw = ref.distribute(pos, cuda(dv))
    nv = dl.nr_lattice_vertices()
    assert nv == b["nv"]
    ks, o2n, _ = canonical(dl.hash_table().m_keys_tensor[:nv].cpu().numpy())
    rks, ro2n, _ = canonical(ref.table.keys[:ref.nv()].cpu().numpy())
    assert np.array_equal(ks, rks)
    assert np.array_equal(lo.relabel(idx.cpu().numpy(), o2n), lo.relabel(ridx.cpu().numpy(), ro2n))
    assert bits_equal(w.cpu().numpy(), rw.cpu().numpy()) == 0
    assert bits_equal(distributed.cpu().numpy(), rdist.cpu().numpy()) == 0, "distributed rows are not bit-equal"
    assert ours.nr_lattice_vertices() == 0          # the parent lattice stays empty (Lattice.cu:377-390)


@pytest.mark.parametrize("V", [1, 8, 32])
def test_slice_fwd_bwd(built, V):
    b = built
    lv = cases.randn((b["nv"], V), 20 + V)                       # canonical order
    ours, ref = b["ours"].clone_lattice(), b["ref"]
    ours.set_values(cuda(lv[b["o2n"]]))
    sliced = ours.slice_standalone_with_precomputation(b["pos"], b["idx"], b["w"])
    rsliced = ref.slice_with_precomputation(b["pos"], cuda(lv[b["ro2n"]]), b["ridx"], b["rw"])
    assert bits_equal(sliced.cpu().numpy(), rsliced.cpu().numpy()) == 0, "slice forward is not bit-equal (same FMA chain expected)"
    assert_close(sliced.cpu().numpy(), lo.slice_fwd(lv, b["cpu"]["indices"], b["cpu"]["weights"], b["n"]), 1e-6, "slice vs oracle")
    g = cases.randn((b["n"], V), 30 + V)
    ours.slice_backwards_standalone_with_precomputation_no_homogeneous(b["pos"], cuda(g), b["idx"], b["w"])
    grad = ours.values().cpu().numpy()[b["n2o"]]
    rgrad = agg(ref.slice_backwards(cuda(g), b["ridx"], b["rw"]).cpu().numpy(), b["ro2n"])
    assert_close(grad, rgrad, TOL_GRADS, "slice backward vs reference kernels")
    assert_close(grad, lo.slice_bwd(g, b["cpu"]["indices"], b["cpu"]["weights"], b["nv"]), TOL_GRADS, "slice backward vs oracle")


def test_slice_no_precomputation(built):
    b = built
    lv = cases.randn((b["nv"], 8), 28)
    ours, ref = b["ours"].clone_lattice(), b["ref"]
    ours.set_values(cuda(lv[b["o2n"]]))
    for shift in (0.0, 0.013):      # shifted positions: some simplex vertices are missing -> -1 entries
        pos = cuda((b["pos_np"] + np.float32(shift)).astype(np.float32))
        s, i, w = ours.slice_standalone_no_precomputation(pos)
        rs, ri, rw = ref.slice_no_precomputation(pos, cuda(lv[b["ro2n"]]))
        assert np.array_equal(lo.relabel(i.cpu().numpy(), b["o2n"]), lo.relabel(ri.cpu().numpy(), b["ro2n"]))
        assert bits_equal(w.cpu().numpy(), rw.cpu().numpy()) == 0
        assert bits_equal(s.cpu().numpy(), rs.cpu().numpy()) == 0
        if shift:
            assert int((i < 0).sum()) > 0


def test_gather_fwd_bwd(built):
    b = built
    V = 8
    lv = cases.randn((b["nv"], V), 28)
    ours, ref = b["ours"].clone_lattice(), b["ref"]
    ours.set_values(cuda(lv[b["o2n"]]))
    g = ours.gather_standalone_with_precomputation(b["pos"], b["idx"], b["w"])
    rg = ref.gather_with_precomputation(b["pos"], cuda(lv[b["ro2n"]]), b["ridx"], b["rw"])
    assert bits_equal(g.cpu().numpy(), rg.cpu().numpy()) == 0
    assert_close(g.cpu().numpy(), lo.gather_fwd(lv, b["cpu"]["indices"], b["cpu"]["weights"], b["n"]), 1e-6, "gather vs oracle")
    gg = cases.randn((b["n"], (b["d"] + 1) * (V + 1)), 40)
    ours.gather_backwards_standalone_with_precomputation(b["pos"], cuda(gg), b["idx"], b["w"])
    grad = ours.values().cpu().numpy()[b["n2o"]]
    rgrad = agg(ref.gather_backwards(cuda(gg), b["ridx"], b["rw"]).cpu().numpy(), b["ro2n"])
    assert_close(grad, rgrad, TOL_GRADS, "gather backward vs reference kernels")
    assert_close(grad, lo.gather_bwd(gg, b["cpu"]["indices"], b["cpu"]["weights"], b["nv"], V), TOL_GRADS, "gather backward vs oracle")


@pytest.mark.parametrize("V,nc", [(32, 7), (128, 20), (8, 4), (64, 16), (256, 20)])
def test_slice_classify(built, V, nc):
    b = built
    if b["d"] != 3 and (V, nc) not in ((32, 7), (64, 16), (8, 4)):
        pytest.skip("reference slice_classify instantiations for pos_dim 5 are built for three (V, nc) pairs")
    n, d = b["n"], b["d"]
    lv = cases.randn((b["nv"], V), 20 + V)
    dw = (cases.randn((n, d + 1), 50) * 0.05).astype(np.float32)
    cw = (cases.randn((nc, V), 51) * 0.2).astype(np.float32)
    cb = cases.randn((nc,), 52)
    gl = cases.randn((n, nc), 53)
    ours, ref = b["ours"].clone_lattice(), b["ref"]
    lv_ours = cuda(lv[b["o2n"]])
    ours.set_values(lv_ours)
    logits = ours.slice_classify_with_precomputation(b["pos"], cuda(dw), cuda(cw), cuda(cb), nc, b["idx"], b["w"])
    rlogits = ref.slice_classify_with_precomputation(b["pos"], cuda(lv[b["ro2n"]]), cuda(dw), cuda(cw), cuda(cb), b["ridx"], b["rw"])
    exp, _ = lo.slice_classify_fwd(lv, b["cpu"]["indices"], b["cpu"]["weights"], dw, cw, cb, n)
    assert_close(logits.cpu().numpy(), rlogits.cpu().numpy(), TOL_VALUES, "slice_classify logits vs reference kernels")
    assert_close(logits.cpu().numpy(), exp, TOL_VALUES, "slice_classify logits vs oracle")
    g_lv, g_dw, g_w, g_b = [torch.zeros_like(t) for t in (lv_ours, cuda(dw), cuda(cw), cuda(cb))]
    ours.slice_classify_backwards_with_precomputation(cuda(gl), b["pos"], lv_ours, cuda(dw), cuda(cw), cuda(cb), nc,
                                                      g_lv, g_dw, g_w, g_b, b["idx"], b["w"])
    r = ref.slice_classify_backwards(cuda(gl), cuda(lv[b["ro2n"]]), cuda(dw), cuda(cw), cuda(cb), b["ridx"], b["rw"])
    e = lo.slice_classify_bwd(gl, lv, b["cpu"]["indices"], b["cpu"]["weights"], dw, cw, n)
    got = [g_lv.cpu().numpy()[b["n2o"]], g_dw.cpu().numpy(), g_w.cpu().numpy(), g_b.cpu().numpy()]
    refs = [agg(r[0].cpu().numpy(), b["ro2n"]), r[1].cpu().numpy(), r[2].cpu().numpy(), r[3].cpu().numpy()]
    for name, a, rr, ee in zip(("grad_values", "grad_delta_w", "grad_W", "grad_b"), got, refs, e):
        assert_close(a, ee, TOL_GRADS, f"slice_classify {name} vs oracle")
        assert_close(a, rr, TOL_GRADS, f"slice_classify {name} vs reference kernels")


def _canon_rowidx(r, o2n_n, n2o_q):
    # a raw 0 is either vertex id 0 or a cell the reference never writes (zeros buffer, Lattice.cu:600):
    # both sides mark it -3
    out = lo.relabel(r, o2n_n).astype(np.int32)
    out[r == 0] = -3
    return out[n2o_q]


def _rowidx(lat_q, lat_n, dil, flip, o2n_n, n2o_q, F):
    r = lat_q.im2rowindices(lat_n, F, dil, flip).cpu().numpy().reshape(-1, F, lat_n.val_dim())[:, :, 0]
    return _canon_rowidx(r, o2n_n, n2o_q)


def _ref_rowidx(ref_q, ref_n, dil, flip, o2n_n, n2o_q, F):
    r = ref_q.im2rowindices(ref_n, 1, dil, flip).cpu().numpy().reshape(-1, F)
    return _canon_rowidx(r, o2n_n, n2o_q)


def _table(lat_q, lat_n, dil, o2n_n, n2o_q):
    """our neighbour table (ids, -1 absent, -2 not examined) in canonical numbering"""
    t = lat_q._neighbour_table(lat_n, dil).cpu().numpy()
    return lo.relabel(t, o2n_n)[n2o_q]


def test_neighbour_tables_same_level(built):
    b = built
    F = 2 * (b["d"] + 1) + 1
    ours = b["ours"].clone_lattice()
    ours.set_values(torch.zeros((b["nv"], 1), device="cuda"))
    for dil in (1, 2):
        exp = lo.neighbour_table(b["ks"], b["ks"], 0, dil)
        for flip in (False, True):
            got = _rowidx(ours, ours, dil, flip, b["o2n"], b["n2o"], F)
            ref = _ref_rowidx(b["ref"], b["ref"], dil, flip, b["ro2n"], b["rn2o"], F)
            assert ((got == ref) | (got == -3) | (ref == -3)).all(), f"im2rowindices dil={dil} flip={flip} differs from the reference kernels"
        assert np.array_equal(_table(ours, ours, dil, b["o2n"], b["n2o"]), exp), "neighbour table differs from the oracle"


def test_coarse_levels(built):
    b = built
    F = 2 * (b["d"] + 1) + 1
    ours = b["ours"].clone_lattice()
    ours.set_values(torch.zeros((b["nv"], 1), device="cuda"))
    coarse = ours.create_coarse_verts_naive(b["pos"])
    rcoarse = b["ref"].create_coarse_verts_naive(b["pos"])
    nvc = coarse.nr_lattice_vertices()
    cks, co2n, cn2o = canonical(coarse.hash_table().m_keys_tensor[:nvc].cpu().numpy())
    rcks, rco2n, rcn2o = canonical(rcoarse.table.keys[:rcoarse.nv()].cpu().numpy())
    assert np.array_equal(cks, rcks) and nvc == len(rcks)      # (the reference may hold duplicates of some keys)
    assert coarse.lvl() == 2 and coarse.m_sigmas == [s * 2.0 for s in ours.m_sigmas]
    coarse.set_values(torch.zeros((nvc, 1), device="cuda"))
    # coarse <- fine  (coarsen forward) and fine <- coarse (finefy forward / coarsen backward)
    got = _rowidx(coarse, ours, 1, False, b["o2n"], cn2o, F)
    ref = _ref_rowidx(rcoarse, b["ref"], 1, False, b["ro2n"], rcn2o, F)
    # (the vertex with id 0 differs between the two runs, so cells naming it are excluded)
    same = (got == ref) | (got == -3) | (ref == -3)
    assert same.all() and ((got == -3) | (ref == -3)).mean() < 0.5
    assert np.array_equal(_table(coarse, ours, 1, b["o2n"], cn2o), lo.neighbour_table(cks, b["ks"], 1, 1))
    for flip in (False, True):
        got = _rowidx(ours, coarse, 1, flip, co2n, b["n2o"], F)
        ref = _ref_rowidx(b["ref"], rcoarse, 1, flip, rco2n, b["rn2o"], F)
        assert ((got == ref) | (got == -3) | (ref == -3)).all()
    assert np.array_equal(_table(ours, coarse, 1, co2n, b["n2o"]), lo.neighbour_table(b["ks"], cks, -1, 1))
    # coarsen<d> kernel (create_coarse_verts)
    kc = ours.create_coarse_verts()
    rkc = b["ref"].create_coarse_verts()
    kk, _, _ = canonical(kc.hash_table().m_keys_tensor[:kc.nr_lattice_vertices()].cpu().numpy())
    rkk, _, _ = canonical(rkc.table.keys[:rkc.nv()].cpu().numpy())
    assert np.array_equal(kk, rkk)
    assert np.array_equal(kk, lo.coarsen_keys(b["ks"]))


@pytest.mark.parametrize("Cin,Cout", [(8, 16), (32, 32), (3, 5), (64, 128)])
def test_conv_fwd_wgrad_dgrad(built, Cin, Cout):
    b = built
    F = 2 * (b["d"] + 1) + 1
    lv = cases.randn((b["nv"], Cin), 70 + Cin)
    fb = (cases.randn((F * Cin, Cout), 60) * 0.1).astype(np.float32)
    ours = b["ours"].clone_lattice()
    ours.set_values(cuda(lv[b["o2n"]]))
    out = ours.convolve_im2row_standalone(cuda(fb), 1, ours, False)
    got = out.values().cpu().numpy()[b["n2o"]]
    table = lo.neighbour_table(b["ks"], b["ks"], 0, 1)
    exp = lo.conv_fwd(lv, table, fb)
    assert_close(got, exp, TOL_VALUES, "conv forward vs oracle")
    ref = b["ref"]
    if ref.k.has(f"im2row<{b['d']},{Cin}>"):
        rout = first(ref.convolve(cuda(fb), ref, cuda(lv[b["ro2n"]]), 1, False).cpu().numpy(), b["rn2o"])
        assert_close(got, rout, TOL_VALUES, "conv forward vs reference im2row+mm")
        rows = ours.im2row(ours, F, 1, False).cpu().numpy()[b["n2o"]]
        rrows = first(ref.im2row(ref, cuda(lv[b["ro2n"]]), 1, False).cpu().numpy(), b["rn2o"])
        assert bits_equal(rows, rrows) == 0, "im2row differs from the reference"
    # weight gradient and data gradient
    g = cases.randn((b["nv"], Cout), 80 + Cout)
    gw = ours.conv_weight_grad(ours, cuda(g[b["o2n"]]), F, 1).cpu().numpy()
    assert_close(gw, lo.conv_wgrad(lv, table, g), TOL_GRADS, "conv weight gradient vs oracle")
    from lattice_net_b200 import Lattice
    fbw = Lattice.filter_for_data_grad(cuda(fb), F, Cin)
    assert np.array_equal(fbw.cpu().numpy(), lo.filter_for_dgrad(fb, F, Cin, Cout))
    q = ours.clone_lattice()
    q.set_values(cuda(g[b["o2n"]]))
    dg = ours.convolve_im2row_standalone(fbw, 1, q, True).values().cpu().numpy()[b["n2o"]]
    assert_close(dg, lo.conv_fwd(g, table, lo.filter_for_dgrad(fb, F, Cin, Cout), flip=True), TOL_GRADS, "conv data gradient vs oracle")
    # data gradient must be the adjoint of the forward: <conv(x), g> == <x, dgrad(g)>
    lhs = float((exp.astype(np.float64) * g).sum())
    rhs = float((lv.astype(np.float64) * dg).sum())
    assert abs(lhs - rhs) <= 1e-3 * max(abs(lhs), 1.0)


def test_row2im(built):
    b = built
    F = 2 * (b["d"] + 1) + 1
    V = 8
    lv = cases.randn((b["nv"], V), 28)
    ours, ref = b["ours"].clone_lattice(), b["ref"]
    ours.set_values(cuda(lv[b["o2n"]]))
    rows = ours.im2row(ours, F, 1, False)
    back = ours.row2im(rows, 1, F, 16, ours).cpu().numpy()[b["n2o"]]
    rrows = ref.im2row(ref, cuda(lv[b["ro2n"]]), 1, False)
    rback = first(ref.row2im(rrows, ref, V, 1).cpu().numpy(), b["rn2o"])
    assert_close(back, rback, 1e-6, "row2im vs reference kernels")
    table = lo.neighbour_table(b["ks"], b["ks"], 0, 1)
    assert_close(back, lo.row2im(lo.im2row(lv, table), table, V), 1e-6, "row2im vs oracle")


def test_table_overflow_is_reported():
    """The reference spins forever when the table is full (HashTableGPU.cuh:443-484); we must raise."""
    from lattice_net_b200 import Lattice
    from lattice_net_b200._cabi import LatticeBackendError
    lat = Lattice(64, [(0.01, 3)])
    pos = cuda(cases.box_surface(2048, 3))
    lat.begin_splat()
    lat.splat_standalone(pos, torch.zeros((2048, 1), device="cuda"))
    with pytest.raises(LatticeBackendError, match="full"):
        lat.nr_lattice_vertices()


def test_empty_and_tiny_inputs():
    from lattice_net_b200 import Lattice
    lat = Lattice(1000, [(0.05, 3)])
    lat.begin_splat()
    idx, w = lat.splat_standalone(torch.zeros((0, 3), device="cuda"), torch.zeros((0, 2), device="cuda"))
    assert idx.numel() == 0 and lat.nr_lattice_vertices() == 0
    lat.begin_splat()
    idx, w = lat.splat_standalone(torch.tensor([[0.01, 0.02, 0.03]], device="cuda"), torch.ones((1, 2), device="cuda"))
    assert lat.nr_lattice_vertices() == 4 and abs(float(w.sum()) - 1.0) < 1e-6
    assert abs(float(lat.values()[:4].sum()) - 2.0) < 1e-5


def test_duplicate_points_and_crowded_table():
    """Collisions and contention: 1000 copies of one point (every insert after the first finds its key present, the
    accumulation hits four rows 1000 times) mixed with random points, in a table sized for an 85 % load factor (long
    linear probes).  Structure must still equal the oracle's exactly, values within the fp32 tolerance."""
    from lattice_net_b200 import Lattice
    rng = np.random.RandomState(11)
    pos_np = np.concatenate([np.tile(np.array([[0.123, -0.456, 0.789]], np.float32), (1000, 1)),
                             (rng.rand(1000, 3).astype(np.float32) - 0.5)], 0)
    pos_np = np.ascontiguousarray(pos_np[rng.permutation(len(pos_np))])
    sig = [0.05] * 3
    cpu = lo.build_lattice(pos_np, sig)
    capacity = int(cpu["nv"] / 0.85) + 1
    vals_np = cases.randn((len(pos_np), 4), 12)
    lat = Lattice(capacity, [(0.05, 3)])
    lat.begin_splat()
    idx, w = lat.splat_standalone(cuda(pos_np), cuda(vals_np))
    nv = lat.nr_lattice_vertices()                      # raises if the table overflowed
    assert nv == cpu["nv"]
    ks, o2n, n2o = canonical(lat.hash_table().m_keys_tensor[:nv].cpu().numpy())
    assert np.array_equal(ks, cpu["keys"])
    assert len(np.unique(ks, axis=0)) == nv             # no key stored twice
    assert np.array_equal(lo.relabel(idx.cpu().numpy(), o2n), cpu["indices"])
    assert bits_equal(w.cpu().numpy(), cpu["weights"]) == 0
    exp = lo.splat_accumulate(vals_np, cpu["indices"], cpu["weights"], nv)
    assert_close(lat.values()[:nv].cpu().numpy()[n2o], exp, TOL_VALUES, "splat values with 1000 coincident points")
    # the same table answers look-ups for the slice without precomputation
    l2 = lat.clone_lattice()
    lv = cases.randn((nv, 8), 13)
    l2.set_values(cuda(lv[o2n]))
    s, i2, w2 = l2.slice_standalone_no_precomputation(cuda(pos_np))
    assert np.array_equal(lo.relabel(i2.cpu().numpy(), o2n), cpu["indices"])
    assert_close(s.cpu().numpy(), lo.slice_fwd(lv, cpu["indices"], cpu["weights"], len(pos_np)), 1e-6, "slice through a crowded table")


def test_golden_vectors_match_cuda_path(golden_dir):
    """Committed reference outputs (made by oracle/make_golden.py from the reference kernels)."""
    from lattice_net_b200 import Lattice
    found = 0
    for name, spec in cases.CASES.items():
        path = os.path.join(golden_dir, f"{name}.npz")
        if not os.path.isfile(path):
            continue
        found += 1
        g = np.load(path)
        lat = Lattice(int(g["capacity"]), [(float(s), 1) for s in g["sigmas"]])
        lat.begin_splat()
        idx, w = lat.splat_standalone(cuda(g["positions"]), cuda(g["splat_in"]))
        nv = lat.nr_lattice_vertices()
        assert nv == int(g["nv"])
        ks, o2n, n2o = canonical(lat.hash_table().m_keys_tensor[:nv].cpu().numpy())
        assert np.array_equal(ks, g["keys"])
        assert np.array_equal(lo.relabel(idx.cpu().numpy(), o2n), g["indices"])
        assert bits_equal(w.cpu().numpy(), g["weights"]) == 0
        assert_close(lat.values()[:nv].cpu().numpy()[n2o], g["splat_values"], TOL_VALUES, f"{name}: splat values vs golden")
        lat2 = lat.clone_lattice()
        lat2.set_values(cuda(g["lv8"][o2n]))
        F = 2 * (g["positions"].shape[1] + 1) + 1
        conv = lat2.convolve_im2row_standalone(cuda(g["conv_filter"]), 1, lat2, False).values().cpu().numpy()[n2o]
        assert_close(conv, g["conv8_16"], TOL_VALUES, f"{name}: conv vs golden")
        s = lat2.slice_standalone_with_precomputation(cuda(g["positions"]), idx, w)
        assert bits_equal(s.cpu().numpy(), g["slice8"]) == 0
    if found == 0:
        pytest.skip("no golden vectors committed yet")


def _gradient_agreement(named_a, grads_b):
    """Statistics of two gradient sets (name -> ndarray): per-tensor max error relative to the tensor's scale, relative
    L2 error and cosine over everything, median / 99th percentile of the element errors."""
    per_tensor, rel, fa, fb = {}, [], [], []
    num = den = 0.0
    for name, ga in named_a:
        gb = grads_b[name]
        ga, gb = ga.astype(np.float64).ravel(), gb.astype(np.float64).ravel()
        scale = max(np.abs(gb).max(), 1e-30)
        per_tensor[name] = float(np.abs(ga - gb).max() / scale)
        rel.append(np.abs(ga - gb) / scale)
        fa.append(ga)
        fb.append(gb)
        num += float(((ga - gb) ** 2).sum())
        den += float((gb ** 2).sum())
    fa, fb, rel = np.concatenate(fa), np.concatenate(fb), np.concatenate(rel)
    return dict(per_tensor=per_tensor, l2=(num / den) ** 0.5, cosine=float(fa @ fb / (np.linalg.norm(fa) * np.linalg.norm(fb))),
                median=float(np.median(rel)), p99=float(np.quantile(rel, 0.99)))


def _assert_flip_robust(st, what):
    """Bounds that hold whether or not a ReLU / LeakyReLU gate flipped between the two evaluations (measured on the B200,
    scripts/diag_cpu_port.py: flipped attempts reach 0.17 on single tensors, L2 3.5e-3, cosine 0.99999, p99 4.8e-3), and
    that an indexing / algebra bug in any gradient path does not meet."""
    worst = max(st["per_tensor"].items(), key=lambda kv: kv[1])
    assert worst[1] <= 0.5, f"{what}: gradient of {worst[0]} max rel err {worst[1]:.3e}"
    assert st["l2"] <= 2e-2, f"{what}: relative L2 error over all gradients {st['l2']:.3e}"
    assert st["cosine"] >= 0.9995, f"{what}: cosine of the concatenated gradients {st['cosine']:.6f}"
    assert st["median"] <= 2e-3 and st["p99"] <= 3e-2, f"{what}: element errors median {st['median']:.2e} / p99 {st['p99']:.2e}"


def test_lnn_model_matches_cpu_port():
    """Model-level parity: the whole LatticeNet forward + backward through the CUDA path vs the
    torch-CPU re-expression (oracle/cpu_port.py) with the same parameters and the same level-1
    vertex numbering (the model treats vertex 0 specially, lattice_modules.py:72-94).

    Forward: logits to 1e-4, loss to 1e-5 on every evaluation (measured: 4e-6 / 2e-7).
    Backward: two fp32 evaluations of this 40-layer network differ in the last bits of the splatted / scattered values
    (atomic order), which now and then flips a ReLU / LeakyReLU gate whose pre-activation sits within ~1e-7 of zero --
    invisible in the logits, but it moves the gradients of a few tensors by percents (scripts/diag_cpu_port.py on the
    B200: an evaluation agrees either to ~7e-6 on every tensor or has 1..11 of the 154 tensors off by 2e-2..2e-1).
    So: every evaluation must meet flip-robust bounds, and EVERY gradient tensor must agree to the tight per-tensor
    bound in at least one evaluation (evaluations cycle over four clouds, so a tensor is never hostage to one gate) --
    a wrong gradient path fails on all of them.  Runs at the product default: tcgen05 3xTF32 convolutions and 1x1 layers."""
    from lattice_net_b200 import Lattice, ModelParams, lattice as lattice_mod
    from lattice_net_b200.losses import segmentation_loss
    assert lattice_mod.CONV_PRECISION == 1
    from lattice_net_b200.models import LNN
    from oracle import cpu_port
    torch.manual_seed(0)
    dev = torch.device("cuda", 0)
    Lattice(60000, [(0.05, 3)])          # the module constructors read the static expected position dimension
    model = LNN(7, ModelParams(), device=dev)
    labels_np = np.random.RandomState(3).randint(0, 7, 2048)
    labels = cuda(labels_np)
    vals = torch.zeros((2048, 1), device=dev)
    cpu = None
    best = {}
    best_l2 = float("inf")
    for attempt in range(12):
        pos_np = cases.box_surface(2048, (0, 3, 4, 1)[attempt % 4])
        pos = cuda(pos_np)
        lattice = Lattice(60000, [(0.05, 3)])
        for p in model.parameters():
            p.grad = None
        logsm, logits = model(lattice, pos, vals)
        # Gradients are compared under the NLL term alone.  The Lovasz term is piecewise linear in the SORTED errors:
        # two errors that differ in the last bits swap places between two fp32 evaluations and the gradients of those
        # two points jump by O(1/|union|).  Its VALUE is compared below.
        loss = torch.nn.functional.nll_loss(logsm, labels)
        loss.backward()
        full_loss = segmentation_loss(logsm.detach(), labels)
        l1 = model.last_level1_lattice
        keys = l1.hash_table().m_keys_tensor[:l1.nr_lattice_vertices()].cpu().numpy()
        if cpu is None:              # after the first forward: the lazily created layers exist
            cpu = cpu_port.CpuLNN(7, ModelParams())
            cpu.load_state_dict({k: v.detach().cpu() for k, v in model.state_dict().items()}, strict=True)
        for p in cpu.parameters():
            p.grad = None
        clogsm, clogits = cpu(pos_np, torch.zeros(2048, 1), [0.05] * 3, level1_keys=keys)
        closs = torch.nn.functional.nll_loss(clogsm, torch.from_numpy(labels_np))
        closs.backward()
        cfull_loss = segmentation_loss(clogsm.detach(), torch.from_numpy(labels_np))
        assert_close(logits.detach().cpu().numpy(), clogits.detach().numpy(), 1e-4, "LNN logits vs CPU port")
        assert abs(loss.item() - closs.item()) <= 1e-5 * abs(closs.item())
        assert abs(full_loss.item() - cfull_loss.item()) <= 1e-3 * abs(cfull_loss.item()), "0.5 Lovasz + 0.5 NLL value"
        cpu_grads = {n: p.grad.numpy() for n, p in cpu.named_parameters() if p.grad is not None}
        st = _gradient_agreement([(n, p.grad.detach().cpu().numpy()) for n, p in model.named_parameters() if p.grad is not None], cpu_grads)
        assert len(st["per_tensor"]) > 100
        _assert_flip_robust(st, f"LNN gradients vs CPU port (evaluation {attempt})")
        for name, err in st["per_tensor"].items():
            best[name] = min(best.get(name, float("inf")), err)
        best_l2 = min(best_l2, st["l2"])
        # tight bounds (north_star: gradients <= 1e-3): per tensor 1e-3 of its scale in some evaluation, 1e-3 relative L2
        # over ALL gradients in some evaluation
        if max(best.values()) <= TOL_GRADS and best_l2 <= TOL_GRADS:
            break
    worst = max(best.items(), key=lambda kv: kv[1])
    assert worst[1] <= TOL_GRADS, f"gradient of {worst[0]}: best of 12 evaluations has max rel err {worst[1]:.3e} > 1e-3"
    assert best_l2 <= TOL_GRADS, f"relative L2 error over all gradients: best of 12 evaluations {best_l2:.3e} > 1e-3"


@pytest.mark.parametrize("precision,tol", [(1, 2e-5), (2, 5e-3)])
@pytest.mark.parametrize("Cin,Cout", [(32, 32), (64, 128), (128, 96), (32, 256), (96, 7)])
def test_conv_tensor_core(built, Cin, Cout, precision, tol):
    """tcgen05 path: 3xTF32 must stay inside the fp32 parity tolerance (1e-4); 1xTF32 is the stated
    looser mode (10-bit mantissa operands: <= 5e-3 of the output scale)."""
    from lattice_net_b200 import lattice as lattice_mod
    b = built
    F = 2 * (b["d"] + 1) + 1
    lv = cases.randn((b["nv"], Cin), 70 + Cin)
    fb = (cases.randn((F * Cin, Cout), 61) * 0.1).astype(np.float32)
    bias = cases.randn((Cout,), 62)
    ours = b["ours"].clone_lattice()
    ours.set_values(cuda(lv[b["o2n"]]))
    table = lo.neighbour_table(b["ks"], b["ks"], 0, 1)
    try:
        lattice_mod.set_conv_precision(precision)
        for flip in (False, True):
            out = ours.convolve_im2row_standalone(cuda(fb), 1, ours, flip, bias=cuda(bias))
            got = out.values().cpu().numpy()[b["n2o"]]
            exp = lo.conv_fwd(lv, table, fb, flip=flip) + bias
            assert_close(got, exp, tol, f"tensor-core conv precision={precision} flip={flip}")
        # weight gradient on the tensor cores (MN-major operands, reduction over the vertices)
        g = cases.randn((b["nv"], Cout), 90 + Cout)
        query = ours.clone_lattice()
        gw = query.conv_weight_grad(ours, cuda(g[b["o2n"]]), F, 1).cpu().numpy()
        assert_close(gw, lo.conv_wgrad(lv, table, g), tol, f"tensor-core weight gradient precision={precision}")
    finally:
        lattice_mod.set_conv_precision(_DEFAULT_PRECISION)


@pytest.mark.parametrize("Cin,Cout", [(64, 512), (512, 64), (384, 384)])
def test_conv_tensor_core_wide_layers(built, Cin, Cout):
    """Layers wider than one UMMA tile (the 384 / 512-channel levels of the SemanticKITTI architecture) run on the
    tensor cores as 256-column chunks: forward (+bias), weight gradient and the transposed-filter data gradient of
    ln_conv_bwd must all match the oracle inside the fp32 parity tolerance (3xTF32; 5e-5 of the output scale at K up to 13*512)."""
    from lattice_net_b200 import lattice as lattice_mod
    b = built
    if b["name"] not in ("shapenet", "d5"):
        pytest.skip("one 3-D and one 5-D lattice cover the chunking")
    F = 2 * (b["d"] + 1) + 1
    lv = cases.randn((b["nv"], Cin), 170 + Cin)
    fb = (cases.randn((F * Cin, Cout), 161) * 0.05).astype(np.float32)
    bias = cases.randn((Cout,), 162)
    g = cases.randn((b["nv"], Cout), 190 + Cout)
    ours = b["ours"].clone_lattice()
    ours.set_values(cuda(lv[b["o2n"]]))
    table = lo.neighbour_table(b["ks"], b["ks"], 0, 1)
    try:
        lattice_mod.set_conv_precision(1)
        out = ours.convolve_im2row_standalone(cuda(fb), 1, ours, False, bias=cuda(bias))
        got = out.values().cpu().numpy()[b["n2o"]]
        assert_close(got, lo.conv_fwd(lv, table, fb) + bias, 5e-5, "wide tensor-core conv forward")
        query = ours.clone_lattice()
        grad_in, grad_filter = query.conv_backward(ours, cuda(g[b["o2n"]]), cuda(fb), 1)
    finally:
        lattice_mod.set_conv_precision(_DEFAULT_PRECISION)
    assert_close(grad_filter.cpu().numpy(), lo.conv_wgrad(lv, table, g), 5e-5, "wide tensor-core weight gradient")
    exp_dg = lo.conv_fwd(g, table, lo.filter_for_dgrad(fb, F, Cin, Cout), flip=True)
    assert_close(grad_in.cpu().numpy()[b["n2o"]], exp_dg, 5e-5, "wide tensor-core data gradient")


def test_conv_tensor_core_cross_level(built):
    from lattice_net_b200 import lattice as lattice_mod
    b = built
    Cin, Cout = 64, 64
    F = 2 * (b["d"] + 1) + 1
    lv = cases.randn((b["nv"], Cin), 75)
    fb = (cases.randn((F * Cin, Cout), 63) * 0.1).astype(np.float32)
    fine = b["ours"].clone_lattice()
    fine.set_values(cuda(lv[b["o2n"]]))
    coarse = fine.create_coarse_verts_naive(b["pos"])
    nvc = coarse.nr_lattice_vertices()
    cks, co2n, cn2o = canonical(coarse.hash_table().m_keys_tensor[:nvc].cpu().numpy())
    up = lo.neighbour_table(cks, b["ks"], 1, 1)
    try:
        lattice_mod.set_conv_precision(1)
        got = coarse.convolve_im2row_standalone(cuda(fb), 1, fine, False).values().cpu().numpy()[cn2o]
    finally:
        lattice_mod.set_conv_precision(_DEFAULT_PRECISION)
    assert_close(got, lo.conv_fwd(lv, up, fb), 2e-5, "tensor-core coarsen conv")


# nv <= 2048 with 1..8 channels per group: one CTA per group; everything else (scene-sized levels, the 512-channel KITTI
# bottleneck with 16 channels per group, widths that are not a multiple of 32) runs the row-tiled kernels
@pytest.mark.parametrize("nv,C", [(983, 32), (1231, 128), (77, 192), (25, 256), (300, 8), (5000, 96), (34809, 64), (2773, 256),
                                  (673, 512), (57169, 32), (10501, 384), (3001, 48), (20000, 256)])   # the last two of 384 / 256: > 592 row blocks -> pre-reduced partials
@pytest.mark.parametrize("relu", [False, True])
def test_fused_group_norm(nv, C, relu):
    """ln_group_norm_fwd/bwd vs torch.nn.GroupNorm on the [1, C, nv] view the reference uses."""
    from lattice_net_b200.lattice_modules import _GroupNormReLU
    torch.manual_seed(nv + C)
    groups = 32 if C % 32 == 0 else C // 2
    x = torch.randn((nv, C), device="cuda") * 2 + 0.5
    gamma = torch.randn(C, device="cuda")
    beta = torch.randn(C, device="cuda")
    g = torch.randn((nv, C), device="cuda")
    xs = [x.clone().requires_grad_(True) for _ in range(2)]
    ps = [(gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)) for _ in range(2)]
    y0 = torch.nn.functional.group_norm(xs[0].t().unsqueeze(0), groups, ps[0][0], ps[0][1], 1e-5).squeeze(0).t()
    if relu:
        y0 = torch.relu(y0)
    y1 = _GroupNormReLU.apply(xs[1], ps[1][0], ps[1][1], groups, 1e-5, relu)
    assert_close(y1.detach().cpu().numpy(), y0.detach().cpu().numpy(), 1e-5, "fused group norm forward")
    y0.backward(g)
    y1.backward(g)
    assert_close(xs[1].grad.cpu().numpy(), xs[0].grad.cpu().numpy(), 1e-4, "fused group norm dx")
    assert_close(ps[1][0].grad.cpu().numpy(), ps[0][0].grad.cpu().numpy(), 1e-4, "fused group norm dgamma")
    assert_close(ps[1][1].grad.cpu().numpy(), ps[0][1].grad.cpu().numpy(), 1e-4, "fused group norm dbeta")


@pytest.mark.parametrize("precision", [0, 1])
def test_conv_transposed_filter_equals_relayout(built, precision):
    """dgrad with the forward bank read transposed in place == dgrad with the re-laid-out copy."""
    from lattice_net_b200 import Lattice, lattice as lattice_mod
    b = built
    F = 2 * (b["d"] + 1) + 1
    Cin, Cout = 32, 64
    fb = (cases.randn((F * Cin, Cout), 64) * 0.1).astype(np.float32)
    g = cases.randn((b["nv"], Cout), 85)
    lat = b["ours"].clone_lattice()
    lat.set_values(cuda(g))
    try:
        lattice_mod.set_conv_precision(precision)
        a = lat.convolve_im2row_standalone(cuda(fb), 1, lat, True, transposed_filter=True).values()
        fbw = Lattice.filter_for_data_grad(cuda(fb), F, Cin)
        c = lat.convolve_im2row_standalone(fbw, 1, lat, True).values()
    finally:
        lattice_mod.set_conv_precision(_DEFAULT_PRECISION)
    assert tuple(a.shape) == (b["nv"], Cin)
    assert_close(a.cpu().numpy(), c.cpu().numpy(), 1e-6, "transposed-filter dgrad")


# --------------------------------------------------------------------------------------------------
# static-shape mode (fixed rows per level, vertex count on the device) and the CUDA-graph step
def test_static_shape_lattice_matches_dynamic(built):
    """Same cloud through a row-bounded lattice: structure identical, padding rows inert."""
    from lattice_net_b200 import Lattice
    b = built
    nv, d = b["nv"], b["d"]
    bound = -(-int(nv * 1.3) // 128) * 128
    lat = _lattice(b["spec"])
    lat.set_vertex_bounds([bound])
    lat.begin_splat()
    idx, w = lat.splat_standalone(b["pos"], cuda(b["vals_np"]))
    assert lat.nr_lattice_vertices() == bound                 # no device sync: the bound
    assert lat.nr_lattice_vertices_actual() == nv
    ks, o2n, n2o = canonical(lat.hash_table().m_keys_tensor[:nv].cpu().numpy())
    assert np.array_equal(ks, b["cpu"]["keys"])
    assert np.array_equal(lo.relabel(idx.cpu().numpy(), o2n), b["cpu"]["indices"])
    assert bits_equal(w.cpu().numpy(), b["cpu"]["weights"]) == 0
    # convolution over the bounded table: rows < nv equal the oracle, rows >= nv are exactly zero
    F = 2 * (d + 1) + 1
    Cin, Cout = 32, 32
    lv = cases.randn((nv, Cin), 21)
    lv_gpu = np.zeros((bound, Cin), np.float32)
    lv_gpu[n2o_rows(o2n, nv)] = lv                         # canonical row c lives at GPU row n2o[c]
    fb = (cases.randn((F * Cin, Cout), 22) * 0.1).astype(np.float32)
    l2 = lat.clone_lattice()
    l2.set_values(cuda(lv_gpu))
    out = l2.convolve_im2row_standalone(cuda(fb), 1, l2, False).values().cpu().numpy()
    assert out.shape == (bound, Cout)
    table = lo.neighbour_table(b["cpu"]["keys"], b["cpu"]["keys"])
    assert_close(out[n2o_rows(o2n, nv)], lo.conv_fwd(lv, table, fb), TOL_VALUES, "static-shape conv")
    assert not out[nv:].any()


def n2o_rows(o2n, nv):
    """GPU row of each canonical vertex (inverse of old_to_new when there are no duplicate keys)."""
    inv = np.empty(nv, np.int64)
    inv[o2n[:nv]] = np.arange(nv)
    return inv


def test_static_shape_bound_exceeded_is_flagged():
    from lattice_net_b200 import Lattice
    from lattice_net_b200._cabi import LatticeBackendError
    pos = cuda(cases.box_surface(2048, 0))
    lat = Lattice(60000, [(0.05, 3)])
    lat.set_vertex_bounds([128])
    lat.begin_splat()
    idx, w = lat.splat_standalone(pos, torch.zeros((2048, 1), device="cuda"))
    i = idx.cpu().numpy()
    assert i.max() < 128 and (i == -1).any()
    assert np.array_equal(w.cpu().numpy()[i < 0], np.full((i < 0).sum(), -1.0, np.float32))
    with pytest.raises(LatticeBackendError, match="max_vertices"):
        lat.nr_lattice_vertices_actual()


@pytest.mark.parametrize("relu", [False, True])
@pytest.mark.parametrize("nv,rows,C", [(700, 1024, 64), (9000, 12032, 128)])
def test_group_norm_ignores_padding_rows(relu, nv, rows, C):
    from lattice_net_b200.lattice_modules import _GroupNormReLU
    torch.manual_seed(5)
    groups = 32
    x = torch.randn((rows, C), device="cuda") * 3 + 1          # padding rows hold garbage on purpose
    g = torch.randn((rows, C), device="cuda")
    g[nv:] = 0                                                   # upstream gradients of padding rows are zero
    gamma, beta = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
    nv_dev = torch.tensor([nv], dtype=torch.int32, device="cuda")
    xa = x[:nv].clone().requires_grad_(True)
    pa = (gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True))
    ya = torch.nn.functional.group_norm(xa.t().unsqueeze(0), groups, pa[0], pa[1], 1e-5).squeeze(0).t()
    ya = torch.relu(ya) if relu else ya
    ya.backward(g[:nv])
    xb = x.clone().requires_grad_(True)
    pb = (gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True))
    yb = _GroupNormReLU.apply(xb, pb[0], pb[1], groups, 1e-5, relu, nv_dev)
    yb.backward(g)
    assert_close(yb[:nv].detach().cpu().numpy(), ya.detach().cpu().numpy(), 1e-5, "padded group norm forward")
    assert not yb[nv:].detach().cpu().numpy().any() and not xb.grad[nv:].cpu().numpy().any()
    assert_close(xb.grad[:nv].cpu().numpy(), xa.grad.cpu().numpy(), 1e-4, "padded group norm dx")
    assert_close(pb[0].grad.cpu().numpy(), pa[0].grad.cpu().numpy(), 1e-4, "padded group norm dgamma")
    assert_close(pb[1].grad.cpu().numpy(), pa[1].grad.cpu().numpy(), 1e-4, "padded group norm dbeta")


def test_graphed_step_matches_eager_step():
    """One CUDA-graph replay of the whole training step (static-shape lattice) == the eager dynamic-shape step:
    same loss, same parameter gradients, vertex counts reported from the device.

    Which vertex gets id 0 is a race in the hash insert (in the reference too, HashTableGPU.cuh:454) and the
    reference model zeroes that vertex (lattice_modules.py:72-94, 712), so two independent runs of the reference
    model are not comparable; the quirk is switched off here, which makes the model invariant to the numbering.

    Two evaluations of the SAME model on the SAME cloud differ in the last bits of the splatted values (fp32 atomics
    in hash-insertion order), which is enough to flip a ReLU / LeakyReLU gate whose pre-activation sits within
    ~1e-7 of zero.  One flipped gate on a 25..100-vertex coarse level moves single gradient elements by percents
    (scripts/diag_flaky.py, profiles/r01g_gradient_reproducibility.txt: eager-vs-eager runs of clouds 0, 2 and 5
    agree either to ~5e-6 or only to 1e-2..7e-2, never in between; clouds 1, 3 and 4 reproduce to <1e-3 every time).
    The test therefore runs on the reproducible clouds, bounds EVERY attempt by flip-robust statistics
    (_assert_flip_robust), and gives every cloud a few attempts in which every gradient tensor must agree to 2e-3 at
    least once (same function when no gate flips -- an indexing / padding bug never would)."""
    import copy
    from lattice_net_b200 import Lattice, ModelParams, lattice_modules
    from lattice_net_b200.graphed import GraphedTrainStep, estimate_vertex_bounds
    from lattice_net_b200.losses import segmentation_loss
    from lattice_net_b200.models import LNN
    from lattice_net_b200.parallel import GradBucket
    torch.manual_seed(1)
    dev = torch.device("cuda", 0)
    clouds = [(cuda(cases.box_surface(2048, s)), torch.zeros((2048, 1), device=dev),
               cuda(np.random.RandomState(s).randint(0, 7, 2048))) for s in (3, 4, 1)]   # clouds on which repeated eager runs reproduce to ~5e-6 (see below)
    lattice_modules.REFERENCE_VERTEX0_QUIRK = False
    try:
        lat_a = Lattice(60000, [(0.05, 3)])
        model_a = LNN(7, ModelParams(), device=dev)
        with torch.no_grad():
            model_a(lat_a, *clouds[0][:2])                      # creates the lazy parameters
        model_b = copy.deepcopy(model_a)
        lat_b = Lattice(60000, [(0.05, 3)])
        # lr = 0: the replayed optimizer step runs but leaves the parameters where they are
        opt_b = torch.optim.AdamW(model_b.parameters(), lr=0.0, weight_decay=3e-4, amsgrad=True, fused=True, capturable=True)
        bucket_b = GradBucket(model_b.parameters())
        bounds = estimate_vertex_bounds(60000, [(0.05, 3)], [c[0] for c in clouds], 4)
        # gradients are compared under the smooth NLL term (the Lovasz term's gradient jumps when two near-equal
        # errors swap places in its sort, see test_lnn_model_matches_cpu_port)
        nll = torch.nn.functional.nll_loss
        step = GraphedTrainStep(model_b, lat_b, opt_b, nll, 2048, 3, 1, bounds, bucket_b, example=clouds[0])
        # the capture's warm-up passes must leave parameters untouched
        for (na, pa), (nb, pb) in zip(model_a.named_parameters(), model_b.named_parameters()):
            assert torch.equal(pa, pb), f"{na} changed during graph capture"
        replays = 0
        for pos, vals, labels in clouds[1:] + clouds[:1]:
            best = {}
            for attempt in range(6):
                loss_b = step(pos, vals, labels)
                replays += 1
                logsm, _ = model_a(lat_a, pos, vals)
                loss_a = nll(logsm, labels)
                for p in model_a.parameters():
                    p.grad = None
                loss_a.backward()
                torch.cuda.synchronize()
                nv_levels = [l.nr_lattice_vertices() for l in model_a.last_level_lattices]
                assert step.last_vertex_counts() == nv_levels
                assert all(n <= b for n, b in zip(nv_levels, bounds))
                assert abs(loss_a.item() - loss_b.item()) <= 1e-4 * abs(loss_a.item())
                grads_a = {n: p.grad.cpu().numpy() for n, p in model_a.named_parameters() if p.grad is not None}
                names = list(grads_a)
                grads_b = [(n, pb.grad.cpu().numpy()) for (n, _), pb in zip(model_a.named_parameters(), model_b.parameters()) if n in grads_a]
                st = _gradient_agreement(grads_b, grads_a)
                assert len(names) > 100
                _assert_flip_robust(st, f"graphed vs eager gradients (attempt {attempt})")
                for name, err in st["per_tensor"].items():
                    best[name] = min(best.get(name, float("inf")), err)
                if max(best.values()) <= TOL_GRADS:
                    break
            worst = max(best.items(), key=lambda kv: kv[1])
            assert worst[1] <= TOL_GRADS, f"graphed gradient of {worst[0]}: best of 6 attempts has max rel err {worst[1]:.3e} > 1e-3"
        assert step.overflowed_steps() == 0
        steps = {int(st["step"].item()) for st in opt_b.state.values()}
        assert steps == {replays}, "the optimizer step inside the graph did not run once per replay"
    finally:
        lattice_modules.REFERENCE_VERTEX0_QUIRK = True


# --------------------------------------------------------------------------------------------------
# BASELINE configs[2] / [3]: SemanticKITTI- and ScanNet-sized scenes (structure parity + one fwd/bwd pass)
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


def _scene(name):
    spec = SCENES[name]
    if name == "kitti":
        pos = cases.kitti_like(spec["n"], 7)
        vals = np.zeros((spec["n"], 1), np.float32)
    else:
        pos, vals = cases.scannet_like(spec["n"], 8)
    return spec, pos, vals


@pytest.mark.parametrize("name", ["kitti", "scannet"])
def test_scene_sized_structure_and_splat_slice(name):
    """Full-size scans: every level's key set / vertex count equals the oracle's, indices and weights bit-exact,
    and splat -> slice satisfies its size-independent identities (constant field reproduced, linearity)."""
    from lattice_net_b200 import Lattice
    spec, pos_np, vals_np = _scene(name)
    sig = [spec["sigma"]] * 3
    pos = cuda(pos_np)
    lat = Lattice(spec["capacity"], [(spec["sigma"], 3)])
    lat.begin_splat()
    ones = torch.ones((spec["n"], 1), device="cuda")
    idx, w = lat.splat_standalone(pos, ones)
    nv = lat.nr_lattice_vertices()
    cpu = lo.build_lattice(pos_np, sig)
    assert nv == cpu["nv"] and nv < 0.5 * spec["capacity"]
    ks, o2n, n2o = canonical(lat.hash_table().m_keys_tensor[:nv].cpu().numpy())
    assert np.array_equal(ks, cpu["keys"])
    assert np.array_equal(lo.relabel(idx.cpu().numpy(), o2n), cpu["indices"])
    assert bits_equal(w.cpu().numpy(), cpu["weights"]) == 0
    # coarse levels exactly as the model builds them (raw points at 2 sigma, 4 sigma, 8 sigma)
    level = lat
    for lvl in range(1, 4):
        level = level.create_coarse_verts_naive(pos)
        nvc = level.nr_lattice_vertices()
        ck, _, _ = canonical(level.hash_table().m_keys_tensor[:nvc].cpu().numpy())
        cc = lo.build_lattice(pos_np, [s * 2 ** lvl for s in sig])
        assert nvc == cc["nv"] and np.array_equal(ck, cc["keys"])
    # barycentric weights of a point sum to one: slicing the splat of a constant, divided by the splatted mass, is constant
    mass = lat.values()[:nv].contiguous()                       # splat of ones = total weight per vertex
    l2 = lat.clone_lattice()
    l2.set_values(torch.ones((nv, 1), device="cuda"))
    s1 = l2.slice_standalone_with_precomputation(pos, idx, w)
    assert_close(s1.cpu().numpy(), np.ones((spec["n"], 1), np.float32), 1e-5, "slice of a constant field")
    assert abs(mass.sum().item() - spec["n"]) <= 1e-3 * spec["n"]
    # linearity of splat in the values (two independent builds: compare in canonical vertex order)
    v = cuda(cases.randn((spec["n"], 8), 3))

    def splat_canonical(values):
        la = Lattice(spec["capacity"], [(spec["sigma"], 3)])
        la.begin_splat()
        la.splat_standalone(pos, values)
        n_ = la.nr_lattice_vertices()
        _, o2n_, _ = canonical(la.hash_table().m_keys_tensor[:n_].cpu().numpy())
        return agg(la.values()[:n_].cpu().numpy(), o2n_)
    assert_close(splat_canonical(2.0 * v), 2.0 * splat_canonical(v), 1e-5, "splat linearity")


@pytest.mark.parametrize("name", ["kitti", "scannet"])
def test_scene_sized_lnn_forward_backward(name):
    """One LatticeNet training pass at full scan size with the scene's own architecture (tcgen05 convolutions)."""
    from lattice_net_b200 import Lattice, ModelParams, lattice as lattice_mod
    from lattice_net_b200.losses import segmentation_loss
    from lattice_net_b200.models import LNN
    spec, pos_np, vals_np = _scene(name)
    torch.manual_seed(0)
    dev = torch.device("cuda", 0)
    lattice = Lattice(spec["capacity"], [(spec["sigma"], 3)])
    model = LNN(spec["nr_classes"], ModelParams(spec["model"]), device=dev)
    labels = torch.randint(0, spec["nr_classes"], (spec["n"],), device=dev)
    try:
        lattice_mod.set_conv_precision(1)
        logsm, logits = model(lattice, cuda(pos_np), cuda(vals_np))
        loss = segmentation_loss(logsm, labels)
        loss.backward()
    finally:
        lattice_mod.set_conv_precision(_DEFAULT_PRECISION)
    torch.cuda.synchronize()
    assert tuple(logits.shape) == (spec["n"], spec["nr_classes"])
    assert torch.isfinite(loss).item()
    grads = [p.grad for p in model.parameters() if p.grad is not None]
    assert len(grads) > 100 and all(torch.isfinite(g).all().item() for g in grads)
    nvs = [l.nr_lattice_vertices() for l in model.last_level_lattices]
    assert nvs == sorted(nvs, reverse=True) and nvs[0] < 0.5 * spec["capacity"]


# --------------------------------------------------------------------------------------------------
# Round-2 additions: the benched configuration (tcgen05 3xTF32 default, prepared filter slabs, fused epilogues, 1x1
# layers on the convolution kernels, gradients written into the bucket) op by op against the oracle / torch fp32.
def _cross_level_pair(b):
    """(fine handle, coarse handle, canonical maps of the coarse level) of the built cloud."""
    fine = b["ours"].clone_lattice()
    coarse = fine.create_coarse_verts_naive(b["pos"])
    nvc = coarse.nr_lattice_vertices()
    cks, co2n, cn2o = canonical(coarse.hash_table().m_keys_tensor[:nvc].cpu().numpy())
    return fine, coarse, nvc, cks, co2n, cn2o


@pytest.mark.parametrize("Cin,Cout", [(192, 96), (256, 128), (128, 256)])
def test_benched_step_conv_shapes_cross_level(built, Cin, Cout):
    """The finefy / coarsen shapes of the ShapeNet step (Appendix D: 256->128 and 192->96 fine<-coarse, 128->256
    coarse<-fine) at the default precision: forward, weight gradient and the CROSS-LEVEL data gradient vs the oracle."""
    b = built
    if b["name"] == "boundary":
        pytest.skip("two clouds cover the cross-level shapes")
    F = 2 * (b["d"] + 1) + 1
    fine, coarse, nvc, cks, co2n, cn2o = _cross_level_pair(b)
    for query, nbr, nq, nn, q_n2o, n_o2n, n_n2o, qk, nk, lvl_diff in (
            (fine, coarse, b["nv"], nvc, b["n2o"], co2n, cn2o, b["ks"], cks, -1),       # finefy: fine vertices query the coarse level
            (coarse, fine, nvc, b["nv"], cn2o, b["o2n"], b["n2o"], cks, b["ks"], 1)):   # coarsen
        lv = cases.randn((nn, Cin), 300 + Cin + lvl_diff)
        fb = (cases.randn((F * Cin, Cout), 301) * 0.05).astype(np.float32)
        g = cases.randn((nq, Cout), 302 + Cout)
        table = lo.neighbour_table(qk, nk, lvl_diff, 1)
        table_bwd = lo.neighbour_table(nk, qk, -lvl_diff, 1)
        n_h = nbr.clone_lattice()
        n_h.set_values(cuda(lv[n_o2n]))
        q_h = query.clone_lattice()
        got = q_h.convolve_im2row_standalone(cuda(fb), 1, n_h, False).values().cpu().numpy()[q_n2o]
        assert_close(got, lo.conv_fwd(lv, table, fb), 3e-5, f"cross-level conv {Cin}->{Cout} lvl_diff={lvl_diff}")
        g_dev = np.empty_like(g)
        g_dev[q_n2o] = g                                   # canonical -> device order of the query level
        grad_in, grad_filter = q_h.conv_backward(n_h, cuda(g_dev), cuda(fb), 1)
        assert_close(grad_filter.cpu().numpy(), lo.conv_wgrad(lv, table, g), 3e-5, f"cross-level weight gradient lvl_diff={lvl_diff}")
        exp_dg = lo.conv_fwd(g, table_bwd, lo.filter_for_dgrad(fb, F, Cin, Cout), flip=True)
        assert_close(grad_in.cpu().numpy()[n_n2o], exp_dg, 3e-5, f"cross-level data gradient lvl_diff={lvl_diff}")


@pytest.mark.parametrize("Cin,Cout", [(32, 32), (128, 128), (64, 20), (8, 16)])
def test_conv_fused_epilogue_and_prepared_slabs(built, Cin, Cout):
    """bias + skip connection inside the convolution epilogue; a bank prepared ahead of time (one batched launch for many
    banks) gives the same result as the per-call preparation; the module-level Function returns the right gradients for
    the bias and the residual."""
    from lattice_net_b200 import lattice as lattice_mod
    from lattice_net_b200.lattice_funcs import ConvIm2RowLattice
    b = built
    F = 2 * (b["d"] + 1) + 1
    nv = b["nv"]
    lv = cases.randn((nv, Cin), 400 + Cin)
    fb = (cases.randn((F * Cin, Cout), 401) * 0.1).astype(np.float32)
    bias = cases.randn((Cout,), 402)
    res = cases.randn((nv, Cout), 403)
    table = lo.neighbour_table(b["ks"], b["ks"], 0, 1)
    exp = lo.conv_fwd(lv, table, fb) + bias + res
    h = b["ours"].clone_lattice()
    h.set_values(cuda(lv[b["o2n"]]))
    fb_t = cuda(fb)
    out0 = h.convolve_im2row_standalone(fb_t, 1, h, False, bias=cuda(bias), residual=cuda(res[b["o2n"]])).values()
    assert_close(out0.cpu().numpy()[b["n2o"]], exp, 3e-5, "conv + bias + residual (per-call filter preparation)")
    # batched preparation: this bank (both readings) together with two unrelated ones
    others = [cuda((cases.randn((F * 64, 96), 410 + i) * 0.1).astype(np.float32)) for i in range(2)]
    readings = [(fb_t, F, Cin, Cout, False), (fb_t, F, Cout, Cin, True)]
    for o in others:
        readings += [(o, F, 64, 96, False), (o, F, 96, 64, True)]
    launched = lattice_mod.prepare_filters(readings)
    tc = Cin % 32 == 0
    assert launched == (6 if (tc and Cout % 32 == 0) else 5 if tc else 4)
    assert lattice_mod.prepare_filters(readings) == 0, "nothing changed: no launch"
    out1 = h.convolve_im2row_standalone(fb_t, 1, h, False, bias=cuda(bias), residual=cuda(res[b["o2n"]])).values()
    assert_close(out1.cpu().numpy()[b["n2o"]], exp, 3e-5, "conv + bias + residual (prepared slabs)")
    g = cases.randn((nv, Cout), 404)
    q = h.clone_lattice()
    gi1, gf1 = q.conv_backward(h, cuda(g[b["o2n"]]), fb_t, 1)
    assert_close(gf1.cpu().numpy(), lo.conv_wgrad(lv, table, g), 3e-5, "weight gradient")
    assert_close(gi1.cpu().numpy()[b["n2o"]], lo.conv_fwd(g, table, lo.filter_for_dgrad(fb, F, Cin, Cout), flip=True), 3e-5,
                 "data gradient (prepared transposed slabs)")
    fb_t.mul_(2.0)                                                # in-place update: the prepared slabs are stale now
    out2 = h.convolve_im2row_standalone(fb_t, 1, h, False).values()
    assert_close(out2.cpu().numpy()[b["n2o"]], 2.0 * lo.conv_fwd(lv, table, fb), 3e-5, "stale slabs must not be used")
    lattice_mod.invalidate_prepared_filters()
    # autograd: gradients of bias and residual through the Function
    x = cuda(lv[b["o2n"]]).requires_grad_(True)
    w = cuda(fb).requires_grad_(True)
    bi = cuda(bias).requires_grad_(True)
    r = cuda(res[b["o2n"]]).requires_grad_(True)
    y, _ = ConvIm2RowLattice.apply(x, b["ours"].clone_lattice(), w, 1, bi, r)
    gy = cuda(g[b["o2n"]])
    y.backward(gy)
    assert_close(bi.grad.cpu().numpy(), g.sum(0), 1e-5, "bias gradient")
    assert bits_equal(r.grad.cpu().numpy(), gy.cpu().numpy()) == 0
    assert_close(w.grad.cpu().numpy(), lo.conv_wgrad(lv, table, g), 3e-5, "filter gradient through autograd")


@pytest.mark.parametrize("M,K,N", [(1002, 128, 32), (128, 256, 64), (77, 64, 256), (1408, 64, 8), (300, 8, 24), (40000, 256, 64)])
@pytest.mark.parametrize("with_bias,with_res", [(False, False), (True, True)])
def test_linear_on_the_convolution_kernels(M, K, N, with_bias, with_res):
    """The 1x1 layers (lattice_modules.py:806-832: torch.nn.Linear) as filter-extent-1 convolutions: y, dx, dW, db and the
    gradient of the fused skip connection vs torch fp32 (fp64 accumulate on the host for the reference values)."""
    from lattice_net_b200.lattice_modules import linear
    x_np, w_np = cases.randn((M, K), 500 + M), (cases.randn((N, K), 501 + K) * 0.1).astype(np.float32)
    b_np, r_np, g_np = cases.randn((N,), 502), cases.randn((M, N), 503), cases.randn((M, N), 504)
    x, w = cuda(x_np).requires_grad_(True), cuda(w_np).requires_grad_(True)
    bi = cuda(b_np).requires_grad_(True) if with_bias else None
    r = cuda(r_np).requires_grad_(True) if with_res else None
    y = linear(x, w, bi, r)
    y.backward(cuda(g_np))
    x64, w64, g64 = x_np.astype(np.float64), w_np.astype(np.float64), g_np.astype(np.float64)
    exp = x64 @ w64.T + (b_np if with_bias else 0.0) + (r_np if with_res else 0.0)
    assert_close(y.detach().cpu().numpy(), exp, 3e-5, "linear forward")
    assert_close(x.grad.cpu().numpy(), g64 @ w64, 3e-5, "linear dx")
    assert_close(w.grad.cpu().numpy(), g64.T @ x64, 3e-5, "linear dW")
    if with_bias:
        assert_close(bi.grad.cpu().numpy(), g64.sum(0), 1e-5, "linear db")
    if with_res:
        assert bits_equal(r.grad.cpu().numpy(), g_np) == 0


@pytest.mark.parametrize("nv,C", [(983, 32), (77, 192), (5000, 96)])
def test_group_norm_split_sums_the_skip_gradient_in_kernel(nv, C):
    from lattice_net_b200.lattice_modules import _GroupNormReLUSplit
    torch.manual_seed(nv)
    x = torch.randn(nv, C, device="cuda", requires_grad=True)
    gn = torch.nn.GroupNorm(32 if C % 32 == 0 else C // 2, C).cuda()
    with torch.no_grad():
        gn.weight.uniform_(0.5, 1.5)
        gn.bias.uniform_(-0.5, 0.5)
    y, skip = _GroupNormReLUSplit.apply(x, gn.weight, gn.bias, gn.num_groups, gn.eps, True, None)
    gy, gs = torch.randn_like(y), torch.randn_like(y)
    (y * gy).sum().add((skip * gs).sum()).backward()
    got = (x.grad.clone(), gn.weight.grad.clone(), gn.bias.grad.clone())
    x.grad = gn.weight.grad = gn.bias.grad = None
    yr = torch.relu(gn(x.t().unsqueeze(0)).squeeze(0).t())
    ((yr * gy).sum() + (x * gs).sum()).backward()
    assert bits_equal(skip.detach().cpu().numpy(), x.detach().cpu().numpy()) == 0
    assert_close(y.detach().cpu().numpy(), yr.detach().cpu().numpy(), 1e-5, "GN+ReLU forward")
    assert_close(got[0].cpu().numpy(), x.grad.cpu().numpy(), 1e-4, "dx + skip gradient")
    assert_close(got[1].cpu().numpy(), gn.weight.grad.cpu().numpy(), 1e-4, "dgamma")
    assert_close(got[2].cpu().numpy(), gn.bias.grad.cpu().numpy(), 1e-4, "dbeta")


# ---- PointNet pooling glue: the torch_scatter replacements on tie-heavy inputs ------------------
@pytest.mark.parametrize("m,c,nv", [(8192, 64, 1002), (4096, 3, 77), (100, 5, 300), (200000, 16, 30000)])
def test_scatter_max_and_sum_count_vs_torch(m, c, nv):
    """ln_scatter_max == torch_scatter.scatter_max semantics (max per vertex, argmax = a row attaining it, empty vertices
    0 / m), ln_scatter_sum_count == scatter_add + bincount; inputs quantised to a handful of values so ties abound."""
    from lattice_net_b200.lattice_modules import scatter_max, scatter_sum_count
    rng = np.random.RandomState(m + c)
    src_np = (rng.randint(-3, 4, (m, c)) * 0.5).astype(np.float32)          # 7 distinct values: many ties per vertex
    idx_np = rng.randint(0, nv, m).astype(np.int32)
    idx_np[: m // 10] = 0                                                   # a crowded vertex, like the reference's row 0
    src = cuda(src_np).requires_grad_(True)
    idx = cuda(idx_np)
    mx, arg = scatter_max(src, idx, nv)
    exp = np.full((nv, c), -np.inf, np.float32)
    np.maximum.at(exp, idx_np, src_np)
    empty = ~np.isin(np.arange(nv), idx_np)
    exp[empty] = 0.0
    got, arg_np = mx.detach().cpu().numpy(), arg.cpu().numpy()
    assert bits_equal(got, exp) == 0
    assert (arg_np[empty] == m).all()
    rows, cols = np.nonzero(~empty[:, None] & np.ones((1, c), bool))
    a = arg_np[rows, cols]
    assert (a >= 0).all() and (a < m).all()
    assert (idx_np[a] == rows).all(), "argmax row does not belong to the vertex"
    assert bits_equal(src_np[a, cols], exp[rows, cols]) == 0, "argmax row does not attain the maximum"
    # smallest row among the ties (deterministic; torch_scatter leaves the choice open)
    first_best = np.full((nv, c), m, np.int64)
    hit = src_np == exp[idx_np]
    r_idx, c_idx = np.nonzero(hit)
    np.minimum.at(first_best, (idx_np[r_idx], c_idx), r_idx)
    assert np.array_equal(arg_np[~empty], first_best[~empty])
    # backward routes each vertex's gradient to its argmax row
    g = cuda(rng.randn(nv, c).astype(np.float32))
    mx.backward(g)
    exp_g = np.zeros((m, c), np.float32)
    np.add.at(exp_g, (a, cols), g.cpu().numpy()[rows, cols])
    assert_close(src.grad.cpu().numpy(), exp_g, 1e-6, "scatter_max backward")
    sums, counts = scatter_sum_count(src.detach(), idx, nv)
    exp_s = np.zeros((nv, c), np.float64)
    np.add.at(exp_s, idx_np, src_np.astype(np.float64))
    assert_close(sums.cpu().numpy(), exp_s, 1e-5, "scatter sum")
    assert np.array_equal(counts.cpu().numpy(), np.bincount(idx_np, minlength=nv).astype(np.float32))


def test_expand_adds_vertices_and_keeps_the_old_ones(built):
    """Lattice::expand (Lattice.cu:292-348): the expanded lattice keeps every vertex of the original under the SAME id,
    adds the vertices of the noisy copies (their key set = the oracle's over the same expanded positions is not
    reproducible because the noise is drawn on the device, so it is checked structurally), pads the values with zeros."""
    b = built
    lat = b["ours"].clone_lattice()
    nv = b["nv"]
    vals = cuda(cases.randn((nv, 6), 600))
    lat.set_values(vals)
    torch.manual_seed(5)
    exp = lat.expand(b["pos"], 4, 0.01 * b["spec"]["sigmas"][0] * 20, True)
    nv2 = exp.nr_lattice_vertices()
    assert nv2 >= nv and lat.nr_lattice_vertices() == nv, "the original lattice must be left alone"
    k_old = lat.hash_table().m_keys_tensor[:nv].cpu().numpy()
    k_new = exp.hash_table().m_keys_tensor[:nv2].cpu().numpy()
    assert np.array_equal(k_new[:nv], k_old), "original vertices keep their ids"
    assert len(np.unique(k_new, axis=0)) == nv2, "no duplicated vertex"
    assert tuple(exp.values().shape) == (nv2, 6)
    assert bits_equal(exp.values()[:nv].cpu().numpy(), vals.cpu().numpy()) == 0 and float(exp.values()[nv:].abs().max()) == 0.0 if nv2 > nv else True
    # every new vertex is a simplex vertex of some expanded position: noise 0 adds nothing
    same = lat.expand(b["pos"], 2, 0.0, False)
    assert same.nr_lattice_vertices() == nv
    # autograd Function: the gradient of the expanded values is cut back to the original rows (lattice_funcs.py:118-143)
    from lattice_net_b200.lattice_funcs import ExpandLattice
    x = vals.clone().requires_grad_(True)
    ev, wrap = ExpandLattice.apply(x, lat, b["pos"], 3, 0.02, True)
    g = torch.randn_like(ev)
    ev.backward(g)
    assert bits_equal(x.grad.cpu().numpy(), g[:nv].cpu().numpy()) == 0


def test_graphed_step_matches_cpu_port():
    """The BENCHED configuration -- whole step replayed as one CUDA graph on the static-shape lattice, tcgen05 3xTF32
    convolutions, prepared filter slabs, gradients written into the bucket -- against the torch-CPU port: loss to 1e-5,
    every gradient tensor within 1e-3 of its scale in at least one evaluation (gate flips: see
    test_lnn_model_matches_cpu_port), flip-robust bounds on every evaluation."""
    from lattice_net_b200 import Lattice, ModelParams, lattice as lattice_mod
    from lattice_net_b200.graphed import GraphedTrainStep, estimate_vertex_bounds
    from lattice_net_b200.models import LNN
    from lattice_net_b200.parallel import GradBucket
    from oracle import cpu_port
    assert lattice_mod.CONV_PRECISION == 1
    torch.manual_seed(2)
    dev = torch.device("cuda", 0)
    seeds = (0, 3, 4, 1)
    clouds_np = [cases.box_surface(2048, s) for s in seeds]
    labels_np = np.random.RandomState(3).randint(0, 7, 2048)
    vals = torch.zeros((2048, 1), device=dev)
    lat = Lattice(60000, [(0.05, 3)])
    model = LNN(7, ModelParams(), device=dev)
    with torch.no_grad():
        model(lat, cuda(clouds_np[0]), vals)
    opt = torch.optim.AdamW(model.parameters(), lr=0.0, weight_decay=0.0, amsgrad=True, fused=True, capturable=True)
    bucket = GradBucket(model.parameters())
    bounds = estimate_vertex_bounds(60000, [(0.05, 3)], [cuda(c) for c in clouds_np], 4)
    nll = torch.nn.functional.nll_loss
    step = GraphedTrainStep(model, lat, opt, nll, 2048, 3, 1, bounds, bucket, example=(cuda(clouds_np[0]), vals, cuda(labels_np)))
    cpu = cpu_port.CpuLNN(7, ModelParams())
    cpu.load_state_dict({k: v.detach().cpu() for k, v in model.state_dict().items()}, strict=True)
    best, best_l2 = {}, float("inf")
    for attempt in range(12):
        pos_np = clouds_np[attempt % 4]
        loss = step(cuda(pos_np), vals, cuda(labels_np))
        torch.cuda.synchronize()
        nv1 = step.last_vertex_counts()[0]
        keys = model.last_level1_lattice.hash_table().m_keys_tensor[:nv1].cpu().numpy()
        for p in cpu.parameters():
            p.grad = None
        clogsm, _ = cpu(pos_np, torch.zeros(2048, 1), [0.05] * 3, level1_keys=keys)
        closs = nll(clogsm, torch.from_numpy(labels_np))
        closs.backward()
        assert abs(loss.item() - closs.item()) <= 1e-5 * abs(closs.item())
        cpu_grads = {n: p.grad.numpy() for n, p in cpu.named_parameters() if p.grad is not None}
        st = _gradient_agreement([(n, p.grad.detach().cpu().numpy()) for n, p in model.named_parameters()
                                  if p.grad is not None and n in cpu_grads], cpu_grads)
        assert len(st["per_tensor"]) > 100
        _assert_flip_robust(st, f"graphed step vs CPU port (evaluation {attempt})")
        for name, err in st["per_tensor"].items():
            best[name] = min(best.get(name, float("inf")), err)
        best_l2 = min(best_l2, st["l2"])
        if max(best.values()) <= TOL_GRADS and best_l2 <= TOL_GRADS:
            break
    worst = max(best.items(), key=lambda kv: kv[1])
    assert worst[1] <= TOL_GRADS, f"graphed gradient of {worst[0]}: best of 12 evaluations has max rel err {worst[1]:.3e} > 1e-3"
    assert best_l2 <= TOL_GRADS
    assert step.overflowed_steps() == 0
    # the gradients of the lattice operators were written straight into the bucket: most .grad tensors alias their slice
    aliased = sum(1 for p, v in zip(bucket.params, bucket.views) if p.grad is not None and p.grad.data_ptr() == v.data_ptr())
    assert aliased >= 100, f"only {aliased} gradients alias the bucket"


@pytest.mark.parametrize("name", ["kitti", "scannet"])
def test_scene_architecture_matches_cpu_port(name):
    """The SemanticKITTI / ScanNet architectures (64..512-channel levels: chunked tcgen05 convolutions, row-tiled
    GroupNorm for 16 channels per group) on a sub-sampled scan (the CPU port needs minutes at full size): logits to 1e-4,
    loss to 1e-5, gradients within the flip-robust bounds and 1e-3 relative L2 over all tensors in some evaluation."""
    from lattice_net_b200 import Lattice, ModelParams, lattice as lattice_mod
    from lattice_net_b200.models import LNN
    from oracle import cpu_port
    assert lattice_mod.CONV_PRECISION == 1
    spec, pos_full, vals_full = _scene(name)
    torch.manual_seed(0)
    dev = torch.device("cuda", 0)
    n = 6000
    sigma = spec["sigma"] * 2.0          # fewer points: a coarser lattice keeps several points per vertex
    mp = ModelParams(spec["model"])
    Lattice(spec["capacity"], [(sigma, 3)])
    model = LNN(spec["nr_classes"], mp, device=dev)
    cpu = None
    best_l2 = float("inf")
    for attempt in range(4):
        sel = np.random.RandomState(40 + attempt).choice(spec["n"], n, replace=False)
        pos_np, vals_np = np.ascontiguousarray(pos_full[sel]), np.ascontiguousarray(vals_full[sel])
        labels_np = np.random.RandomState(41 + attempt).randint(0, spec["nr_classes"], n)
        lattice = Lattice(spec["capacity"], [(sigma, 3)])
        for p in model.parameters():
            p.grad = None
        logsm, logits = model(lattice, cuda(pos_np), cuda(vals_np))
        loss = torch.nn.functional.nll_loss(logsm, cuda(labels_np))
        loss.backward()
        l1 = model.last_level1_lattice
        keys = l1.hash_table().m_keys_tensor[:l1.nr_lattice_vertices()].cpu().numpy()
        if cpu is None:
            cpu = cpu_port.CpuLNN(spec["nr_classes"], mp, val_dim=spec["val_dim"])
            cpu.load_state_dict({k: v.detach().cpu() for k, v in model.state_dict().items()}, strict=True)
        for p in cpu.parameters():
            p.grad = None
        clogsm, clogits = cpu(pos_np, torch.from_numpy(vals_np), [sigma] * 3, level1_keys=keys)
        closs = torch.nn.functional.nll_loss(clogsm, torch.from_numpy(labels_np))
        closs.backward()
        assert_close(logits.detach().cpu().numpy(), clogits.detach().numpy(), 1e-4, f"{name} architecture: logits vs CPU port")
        assert abs(loss.item() - closs.item()) <= 1e-5 * abs(closs.item())
        cpu_grads = {k: p.grad.numpy() for k, p in cpu.named_parameters() if p.grad is not None}
        st = _gradient_agreement([(k, p.grad.detach().cpu().numpy()) for k, p in model.named_parameters() if p.grad is not None], cpu_grads)
        assert len(st["per_tensor"]) > 100
        _assert_flip_robust(st, f"{name} architecture gradients vs CPU port (evaluation {attempt})")
        best_l2 = min(best_l2, st["l2"])
        if best_l2 <= TOL_GRADS:
            break
    assert best_l2 <= TOL_GRADS, f"{name} architecture: relative L2 error over all gradients, best of 4: {best_l2:.3e}"


# ---- loss and optimizer kernels ------------------------------------------------------------------
@pytest.mark.parametrize("n,nc,ignore", [(2048, 7, -100), (2000, 7, 2), (8192, 20, 0), (1, 3, -100), (1500, 50, -100)])
def test_fused_segmentation_loss_vs_reference_formulation(n, nc, ignore):
    """ln_seg_loss_fwd / _bwd against the per-class loop of the reference's LovaszSoftmax (lovasz_loss.py:41-72) plus
    torch NLL, value (1e-5) and gradient w.r.t. the logits (1e-4 of the gradient scale; ties in the sort are broken
    differently but errors of random logits do not tie)."""
    from lattice_net_b200.losses import _SegLossFn, lovasz_softmax_loop, segmentation_loss
    torch.manual_seed(n + nc)
    logits = (torch.randn(n, nc, device="cuda") * 2.0).requires_grad_(True)
    labels = torch.randint(0, max(nc - 2, 1), (n,), device="cuda")            # the last classes are absent
    lsm = torch.log_softmax(logits, 1)
    loss = segmentation_loss(lsm, labels, ignore)
    assert isinstance(loss.grad_fn, _SegLossFn._backward_cls) or n > 8192
    g, = torch.autograd.grad(loss * 3.0, logits)                               # upstream gradient != 1
    logits64 = logits.detach().double().cpu().requires_grad_(True)
    lsm64 = torch.log_softmax(logits64, 1)
    lab = labels.cpu()
    nll = torch.nn.functional.nll_loss(lsm64, lab, ignore_index=ignore)
    lov = lovasz_softmax_loop(lsm64.exp(), lab, ignore if ignore >= 0 else None) if n > 1 or True else 0.0
    ref = 0.5 * lov + 0.5 * nll
    gr, = torch.autograd.grad(ref * 3.0, logits64)
    assert abs(loss.item() - ref.item()) <= 1e-5 * max(abs(ref.item()), 1e-3), (loss.item(), ref.item())
    assert_close(g.cpu().numpy(), gr.numpy(), 1e-4, "fused loss gradient")


def test_flat_adamw_matches_torch_adamw_amsgrad():
    """optim.FlatAdamW == torch.optim.AdamW(amsgrad=True) (ln_train.py:163-165) over several steps, including a skipped one."""
    from lattice_net_b200.optim import FlatAdamW
    from lattice_net_b200.parallel import GradBucket
    torch.manual_seed(0)
    shapes = [(288, 32), (32,), (7, 128), (1,), (1152, 128), (5, 3, 2)]
    pa = [torch.nn.Parameter(torch.randn(*s, device="cuda")) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    ref = torch.optim.AdamW(pa, lr=1e-3, weight_decay=3e-4, amsgrad=True)
    bucket = GradBucket(pb)
    opt = FlatAdamW(bucket, lr=1e-3, weight_decay=3e-4)
    skip = torch.zeros((), device="cuda")
    opt.found_inf = skip
    for it in range(6):
        grads = [torch.randn(*s, device="cuda") * (0.1 + it) for s in shapes]
        for p, q, g in zip(pa, pb, grads):
            p.grad = g.clone()
            q.grad = g.clone()
        bucket.pack()
        if it == 3:
            skip.fill_(1.0)
            before = [q.detach().clone() for q in pb]
            opt.step()
            skip.zero_()
            assert all(torch.equal(a, q.detach()) for a, q in zip(before, pb)) and opt.steps_taken() == 3
            continue
        ref.step()
        opt.step()
        for p, q in zip(pa, pb):
            assert_close(q.detach().cpu().numpy(), p.detach().cpu().numpy(), 2e-6, f"parameters after step {it}")
    assert opt.steps_taken() == 5
    # 1/world_size folded into the update
    pc = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    pd = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    r2 = torch.optim.AdamW(pc, lr=1e-3, weight_decay=3e-4, amsgrad=True)
    b2 = GradBucket(pd)
    o2 = FlatAdamW(b2, lr=1e-3, weight_decay=3e-4)
    for p, q, s in zip(pc, pd, shapes):
        g = torch.randn(*s, device="cuda")
        p.grad = g * 0.25
        q.grad = g.clone()
    b2.pack()
    r2.step()
    o2.step(grad_scale=0.25)
    for p, q in zip(pc, pd):
        assert_close(q.detach().cpu().numpy(), p.detach().cpu().numpy(), 2e-6, "grad_scale")


@pytest.mark.parametrize("quirk", [True, False])
@pytest.mark.parametrize("val_dim", [1, 4])
def test_fused_pointnet_matches_the_module_by_module_path(built, val_dim, quirk):
    """csrc/ln_pointnet.cu (no per-row tensors) against the reference-shaped path of this repo -- distribute rows, scatter
    mean, torch MLP with weight norm, scatter max, index_select, masks (lattice_modules.py:52-96, 620-733) -- on the SAME
    lattice structure: pooled features + the convolution after them to 1e-5, every parameter gradient to 1e-4."""
    from lattice_net_b200 import lattice_modules as lm
    b = built
    if b["name"] == "boundary":
        pytest.skip("two clouds cover it")
    d = b["d"]
    torch.manual_seed(11)
    vals = cuda(cases.randn((b["n"], val_dim), 700 + val_dim))
    saved = lm.REFERENCE_VERTEX0_QUIRK
    lm.REFERENCE_VERTEX0_QUIRK = quirk
    try:
        lat = _lattice(b["spec"])
        with torch.no_grad():
            dl, distributed, idx, w = lm.DistributeLatticeModule()(lat, b["pos"], vals)
        pn = lm.PointNetModule([16, 32, 64], 32, device=torch.device("cuda", 0))
        assert pn.fused_supported(d, val_dim)
        with torch.no_grad():
            pn._init_layers(d + val_dim, b["pos"].device)
            for layer in pn.layers:                      # non-trivial gains / biases
                layer.weight_g.mul_(torch.rand_like(layer.weight_g) + 0.5)
                layer.bias.uniform_(-0.2, 0.2)
        nv = dl.nr_lattice_vertices()
        g = torch.randn((nv, 32), device="cuda")
        results = []
        for fused in (False, True):
            for p in pn.parameters():
                p.grad = None
            handle = dl.clone_lattice()
            lv, _ = pn.forward_fused(handle, b["pos"], vals, idx, w) if fused else pn(handle, distributed.clone(), idx)
            pooled = handle.values() if not fused else None
            (lv * g).sum().backward()
            results.append((lv.detach().cpu().numpy(), {n_: p.grad.detach().cpu().numpy().copy() for n_, p in pn.named_parameters()}))
        assert_close(results[1][0], results[0][0], 1e-5, "PointNet + first convolution output")
        for name, ga in results[0][1].items():
            assert_close(results[1][1][name], ga, 1e-4, f"fused PointNet gradient of {name}")
    finally:
        lm.REFERENCE_VERTEX0_QUIRK = saved


def test_programmatic_dependent_launch_does_not_change_results(built):
    """ln_set_programmatic_launch(0 / 1): the same chain of dependent kernels (GroupNorm -> convolution -> GroupNorm ->
    slice, forward and backward) with and without programmatic dependent launch; deterministic kernels bit-equal, the
    reduction-based ones within their usual tolerance.  Guards the rule that no kernel touches global memory before its
    griddepcontrol.wait."""
    from lattice_net_b200 import _cabi
    from lattice_net_b200.lattice_funcs import ConvIm2RowLattice, SliceLattice
    from lattice_net_b200.lattice_modules import _GroupNormReLU
    b = built
    F = 2 * (b["d"] + 1) + 1
    C = 64
    torch.manual_seed(3)
    x0 = torch.randn((b["nv"], C), device="cuda")
    w0 = torch.randn((F * C, C), device="cuda") * 0.05
    gam, bet = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.1
    gy = torch.randn((b["n"], C), device="cuda")
    lib = _cabi.load()

    def run():
        x = x0.clone().requires_grad_(True)
        w = w0.clone().requires_grad_(True)
        g1, b1 = gam.clone().requires_grad_(True), bet.clone().requires_grad_(True)
        h = x
        for _ in range(6):      # a dependent chain long enough for several kernels to be in flight; no ReLU: a gate within an ulp of
            # zero would flip with the order of the split-K reductions and hide what this test is after
            h = _GroupNormReLU.apply(h, g1, b1, 32, 1e-5, False, None)
            h, _ = ConvIm2RowLattice.apply(h, b["ours"].clone_lattice(), w, 1)
        s = SliceLattice.apply(h, b["ours"].clone_lattice(), b["pos"], b["idx"], b["w"])
        (s * gy).sum().backward()
        torch.cuda.synchronize()
        return [t.detach().cpu().numpy() for t in (s, x.grad, w.grad, g1.grad, b1.grad)]

    prev = lib.ln_set_programmatic_launch(1)
    try:
        on = run()
        lib.ln_set_programmatic_launch(0)
        off = run()
    finally:
        lib.ln_set_programmatic_launch(prev)
    for a, c, what in zip(on, off, ("sliced", "dx", "dw", "dgamma", "dbeta")):
        assert_close(a, c, 1e-4, f"programmatic launch on vs off: {what}")


def test_fused_delta_weights_match_the_torch_sequence(built):
    """ln_deltaw_fwd / _bwd against gather -> view -> max -> affine -> subtract -> Linear(9 -> 1)
    (lattice_modules.py:465-567) through the slice head module with the fusion switched off: logits to 1e-5, every
    gradient (head parameters and incoming lattice values) to 1e-4."""
    from lattice_net_b200 import lattice_modules as lm
    b = built
    if b["name"] == "boundary":
        pytest.skip("two clouds cover it")
    torch.manual_seed(21)
    dev = torch.device("cuda", 0)
    C, nc = 64, 7
    head = lm.SliceFastCUDALatticeModule(C, nc, 0.0, "none", device=dev)
    lv0 = torch.randn((b["nv"], C), device=dev)
    g = torch.randn((b["n"], nc), device=dev)
    with torch.no_grad():
        head(lv0, b["ours"].clone_lattice(), b["pos"], b["idx"], b["w"])         # lazy layers
        head.linear_deltaW.weight.normal_(0.0, 0.5)
        head.linear_deltaW.bias.fill_(0.1)
        head.gamma.uniform_(0.5, 1.5)
        head.beta.uniform_(-0.3, 0.3)
    outs = []
    for fused in (False, True):
        head.fused_delta_weights = fused
        for p in head.parameters():
            p.grad = None
        lv = lv0.clone().requires_grad_(True)
        logits = head(lv, b["ours"].clone_lattice(), b["pos"], b["idx"], b["w"])
        (logits * g).sum().backward()
        outs.append((logits.detach().cpu().numpy(), lv.grad.cpu().numpy(), {n_: p.grad.cpu().numpy().copy() for n_, p in head.named_parameters()}))
    assert_close(outs[1][0], outs[0][0], 1e-5, "logits with fused delta weights")
    assert_close(outs[1][1], outs[0][1], 1e-4, "gradient of the lattice values")
    for name, ga in outs[0][2].items():
        assert_close(outs[1][2][name], ga, 1e-4, f"gradient of {name}")


@pytest.mark.parametrize("rows,cols,g_dim", [(1152, 32, 1), (16, 4, 0), (64, 32, 0), (2304, 64, 1)])
def test_weight_norm_kernels_vs_torch(rows, cols, g_dim):
    from lattice_net_b200.lattice_modules import _WeightNormFn
    torch.manual_seed(rows + cols)
    v = torch.randn(rows, cols, device="cuda", requires_grad=True)
    gshape = (1, cols) if g_dim == 1 else (rows, 1)
    g = (torch.rand(*gshape, device="cuda") + 0.5).requires_grad_(True)
    up = torch.randn(rows, cols, device="cuda")
    w = _WeightNormFn.apply(v, g, g_dim)
    w.backward(up)
    got = (w.detach().clone(), v.grad.clone(), g.grad.clone())
    v.grad = g.grad = None
    wr = v * (g / v.norm())
    wr.backward(up)
    assert_close(got[0].cpu().numpy(), wr.detach().cpu().numpy(), 1e-6, "weight norm forward")
    assert_close(got[1].cpu().numpy(), v.grad.cpu().numpy(), 1e-5, "weight norm dv")
    assert_close(got[2].cpu().numpy(), g.grad.cpu().numpy(), 1e-5, "weight norm dg")


def test_reference_arm_drives_the_reference_python_unmodified():
    """`bench.py --impl reference`: baseline/_ref holds byte-for-byte copies of the reference's Python layer (sha256
    manifest), the arm imports THEM (not this repo's modules) and two training steps run on the reference kernels."""
    import hashlib
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    manifest_path = os.path.join(root, "baseline", "_ref", "MANIFEST.json")
    if not os.path.isfile(manifest_path):
        pytest.skip("baseline/_ref not installed (python -m oracle.install_ref_py where /root/reference is mounted)")
    _ref()
    with open(manifest_path) as f:
        manifest = json.load(f)["sha256"]
    for rel, digest in manifest.items():
        with open(os.path.join(root, "baseline", "_ref", "latticenet_py", rel), "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == digest, f"{rel} was modified after installation"
    from lattice_net_b200.params import ModelParams
    from oracle import ref_arm
    model = ref_arm.build_reference_model(7, ModelParams())
    assert type(model).__module__ == "latticenet_py.lattice.models"
    assert os.path.join("baseline", "_ref") in sys.modules["latticenet_py.lattice.lattice_modules"].__file__
    pos = cuda(cases.box_surface(2048, 0))
    vals = torch.zeros((2048, 1), device="cuda")
    labels = torch.randint(0, 7, (2048,), device="cuda")
    lat = ref_arm.RefHandle(60000, [0.05] * 3)
    losses = []
    for _ in range(2):
        logsm, logits = model(lat, pos, vals)
        loss = torch.nn.functional.nll_loss(logsm, labels)
        for p in model.parameters():
            p.grad = None
        loss.backward()
        losses.append(loss.item())
    assert all(np.isfinite(losses)) and tuple(logits.shape) == (2048, 7)
    grads = [p.grad for p in model.parameters() if p.grad is not None]
    assert len(grads) > 100 and all(torch.isfinite(g).all().item() for g in grads)
    # no kernel of this library ran on that path
    from lattice_net_b200 import _cabi
    before = _cabi.launch_count()
    model(lat, pos, vals)
    assert _cabi.launch_count() == before


def test_input_path_prepare_cloud_and_pinned_feeder(tmp_path):
    """prepare_cloud's position / value modes (models.py:18-66), the SemanticKITTI file formats, and the
    double-buffered pinned feeder delivering every cloud intact while the previous one is being consumed."""
    from lattice_net_b200.data import PinnedCloudFeeder, SyntheticCloud, prepare_cloud, read_semantic_kitti_scan, write_label_file
    from lattice_net_b200.params import ModelParams
    c = SyntheticCloud(500, 7, 3, with_colour=True, with_intensity=True)
    for pm, vm, pd, vd in [("xyz", "none", 3, 1), ("xyz+rgb", "intensity", 6, 1), ("xyz+intensity", "rgb+height", 4, 4), ("xyz", "rgb+xyz", 3, 6),
                           ("xyz", "height", 3, 1), ("xyz", "xyz", 3, 3), ("xyz", "rgb", 3, 3)]:
        pos, vals, tgt = prepare_cloud(c, ModelParams(dict(positions_mode=pm, values_mode=vm)))
        assert tuple(pos.shape) == (500, pd) and tuple(vals.shape) == (500, vd) and tgt.dtype == torch.int64 and tuple(tgt.shape) == (500,)
        assert pos.is_cuda and pos.is_contiguous() and vals.is_contiguous()
        assert bits_equal(pos[:, :3].cpu().numpy(), c.V) == 0
    pos, vals, _ = prepare_cloud(c, ModelParams(dict(positions_mode="xyz", values_mode="rgb+height")))
    assert bits_equal(vals[:, 3].cpu().numpy(), c.V[:, 1]) == 0 and bits_equal(vals[:, :3].cpu().numpy(), c.C) == 0
    # SemanticKITTI formats
    scan = np.concatenate([c.V, c.I], 1).astype(np.float32)
    scan.tofile(tmp_path / "000000.bin")
    (c.L_gt.reshape(-1).astype(np.uint32) | np.uint32(5 << 16)).tofile(tmp_path / "000000.label")     # upper 16 bits: instance id
    k = read_semantic_kitti_scan(str(tmp_path / "000000.bin"), str(tmp_path / "000000.label"))
    assert bits_equal(k.V, c.V) == 0 and np.array_equal(k.L_gt, c.L_gt)
    logsm = torch.log_softmax(torch.randn(500, 7, device="cuda"), 1)
    written = write_label_file(logsm, str(tmp_path / "pred" / "000000.label"))
    assert np.array_equal(np.fromfile(tmp_path / "pred" / "000000.label", dtype=np.uint32), logsm.argmax(1).cpu().numpy().astype(np.uint32))
    assert written.dtype == np.uint32
    # feeder: 7 clouds through 2 slots, consumed by a slow kernel so that staging really overlaps
    n = 4096
    clouds = [(np.random.RandomState(i).rand(n, 3).astype(np.float32), np.full((n, 1), i, np.float32), np.full((n,), i, np.int64)) for i in range(7)]
    feeder = PinnedCloudFeeder(n, 3, 1, torch.device("cuda", 0))
    feeder.stage(*clouds[0])
    sink = torch.zeros((2048, 2048), device="cuda")
    seen = []
    for i in range(7):
        slot, (p, v, l) = feeder.current()
        if i + 1 < 7:
            feeder.stage(*clouds[i + 1])
        sink = sink @ sink                                   # keeps the compute stream busy
        seen.append((p.clone(), v.clone(), l.clone()))
        feeder.release(slot)
    torch.cuda.synchronize()
    for i, (p, v, l) in enumerate(seen):
        assert bits_equal(p.cpu().numpy(), clouds[i][0]) == 0 and float(v.min()) == float(v.max()) == float(i) and int(l[0]) == i
