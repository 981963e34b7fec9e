"""The oracle checked against itself (properties) and against the golden vectors produced by the
reference's own kernels (tests/golden/*.npz, oracle/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from oracle import cases, lattice_oracle as lo
from tests.util import assert_close, bits_equal

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gold(name):
    path = os.path.join(GOLD, f"{name}.npz")
    if not os.path.isfile(path):
        pytest.skip(f"golden vector {name}.npz not committed yet")
    return np.load(path)


@pytest.mark.parametrize("name", list(cases.CASES))
def test_structure_properties(name):
    spec = cases.CASES[name]
    pos = spec["make"]()
    L = lo.build_lattice(pos, spec["sigmas"])
    n, d = pos.shape
    w = L["weights"].reshape(n, d + 1)
    assert np.allclose(w.sum(1), 1.0, atol=1e-5) and w.min() > -1e-5           # barycentric coordinates
    full = np.concatenate([L["keys"], -L["keys"].sum(1, keepdims=True)], 1)
    assert np.all((full - full[:, :1]) % (d + 1) == 0)                           # all coords congruent mod d+1
    assert len(np.unique(L["keys"], axis=0)) == L["nv"]
    # the simplex vertices of a point are pairwise 1-hop neighbours with remainders 0..d
    sk = L["simplex_keys"][0]
    assert sorted((sk[:, 0] - sk[0, 0]) % (d + 1)) == list(range(d + 1))
    nf, chain = lo.hash_chain_stats(L["keys"], spec["capacity"])
    assert nf == L["nv"] and chain < 300                                          # retrieve()'s probe cap never bites


def test_hash_matches_reference_formula():
    keys = np.array([[1, 2, 3], [-5, 7, 0], [100000, -99999, 4]], np.int32)
    exp = []
    for k in keys:
        h = 0
        for c in k:
            h = (h + int(c)) & 0xFFFFFFFF
            h = (h * 2531011) & 0xFFFFFFFF
        exp.append(h)
    assert list(lo.key_hash(keys)) == exp


def test_neighbour_table_symmetry_and_row2im_adjoint():
    L = lo.build_lattice(cases.box_surface(512, 5), [0.05] * 3)
    T = lo.neighbour_table(L["keys"], L["keys"], 0, 1)
    nv, F = T.shape
    for q in range(0, nv, 7):                       # np of q <-> q is nm of that neighbour
        for a in range(4):
            j = T[q, 2 * a]
            if j >= 0:
                assert T[j, 2 * a + 1] == q
    assert np.array_equal(T[:, F - 1], np.arange(nv))
    rng = np.random.RandomState(0)
    x = rng.randn(nv, 4).astype(np.float32)
    y = rng.randn(nv, F * 4).astype(np.float32)
    lhs = float((lo.im2row(x, T).astype(np.float64) * y).sum())                   # <im2row(x), y>
    rhs = float((x.astype(np.float64) * lo.row2im(y, T, 4)).sum())               # == <x, row2im(y)>
    assert abs(lhs - rhs) < 1e-3 * abs(lhs)


def test_conv_dgrad_is_adjoint_and_wgrad_matches_finite_difference():
    L = lo.build_lattice(cases.box_surface(256, 6), [0.05] * 3)
    T = lo.neighbour_table(L["keys"], L["keys"], 0, 1)
    nv, F = T.shape
    rng = np.random.RandomState(1)
    x = rng.randn(nv, 3).astype(np.float32)
    fb = rng.randn(F * 3, 5).astype(np.float32)
    g = rng.randn(nv, 5).astype(np.float32)
    out = lo.conv_fwd(x, T, fb)
    dg = lo.conv_fwd(g, T, lo.filter_for_dgrad(fb, F, 3, 5), flip=True)
    assert abs(float((out * g).sum()) - float((x * dg).sum())) < 1e-3 * abs(float((out * g).sum()))
    gw = lo.conv_wgrad(x, T, g)
    fb2 = fb.copy()
    fb2[7, 2] += 1e-2
    fd = (float((lo.conv_fwd(x, T, fb2).astype(np.float64) * g).sum()) - float((out.astype(np.float64) * g).sum())) / 1e-2
    assert abs(fd - gw[7, 2]) < 2e-2 * max(abs(fd), 1.0)


def test_splat_slice_gather_backwards_are_the_adjoints_of_their_forwards():
    """The oracle's backward restatements against its forward ones: splat / slice backward is the transpose of
    slice forward, gather backward the transpose of gather forward on the value columns."""
    pos = cases.box_surface(300, 9)
    L = lo.build_lattice(pos, [0.05] * 3)
    n, nv = len(pos), L["nv"]
    rng = np.random.RandomState(2)
    V = 5
    lv = rng.randn(nv, V).astype(np.float32)
    g = rng.randn(n, V).astype(np.float32)
    lhs = float((lo.slice_fwd(lv, L["indices"], L["weights"], n).astype(np.float64) * g).sum())
    rhs = float((lv.astype(np.float64) * lo.slice_bwd(g, L["indices"], L["weights"], nv)).sum())
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0)
    assert np.array_equal(lo.slice_bwd(g, L["indices"], L["weights"], nv), lo.splat_accumulate(g, L["indices"], L["weights"], nv))
    gg = rng.randn(n, 4 * (V + 1)).astype(np.float32)
    gathered = lo.gather_fwd(lv, L["indices"], L["weights"], n)
    gv = gg.reshape(n, 4, V + 1).copy()
    gv[:, :, V] = 0.0                                  # the weight column carries no gradient to the values
    lhs = float((gathered.astype(np.float64) * gv.reshape(n, -1)).sum())
    rhs = float((lv.astype(np.float64) * lo.gather_bwd(gg, L["indices"], L["weights"], nv, V)).sum())
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0)
    # the weight column of the gather is the barycentric weight itself
    assert np.array_equal(gathered.reshape(n, 4, V + 1)[:, :, V].ravel(), L["weights"])


def test_slice_classify_backward_matches_finite_differences():
    """slice_classify_bwd (LatticeGPU.cuh:3628-3756 restated) against central differences of slice_classify_fwd in all
    four of its differentiable inputs: lattice values, delta weights, classifier weight and bias."""
    pos = cases.box_surface(200, 10)
    L = lo.build_lattice(pos, [0.05] * 3)
    n, nv = len(pos), L["nv"]
    rng = np.random.RandomState(4)
    V, nc = 6, 5
    lv = rng.randn(nv, V).astype(np.float32)
    dw = (rng.randn(n, 4) * 0.05).astype(np.float32)
    cw = (rng.randn(nc, V) * 0.3).astype(np.float32)
    cb = rng.randn(nc).astype(np.float32)
    g = rng.randn(n, nc).astype(np.float32)

    def objective(lv_, dw_, cw_, cb_):
        # float64 pre-rounding value: the second return of slice_classify_fwd is the sliced row in float64
        _, s = lo.slice_classify_fwd(lv_, L["indices"], L["weights"], dw_, cw_, cb_, n)
        return float(((s @ np.asarray(cw_, np.float64).T + np.asarray(cb_, np.float64)) * g).sum())

    g_lv, g_dw, g_w, g_b = lo.slice_classify_bwd(g, lv, L["indices"], L["weights"], dw, cw, n)
    eps = 1e-2                                           # exact up to rounding: the objective is bilinear in each input
    for arr, grad, picks in ((lv, g_lv, [(0, 0), (nv // 2, 3), (nv - 1, V - 1)]), (dw, g_dw, [(0, 0), (n // 3, 2), (n - 1, 3)]),
                             (cw, g_w, [(0, 0), (nc - 1, V - 1)]), (cb, g_b, [(0,), (nc - 1,)])):
        for pick in picks:
            plus, minus = arr.copy(), arr.copy()
            plus[pick] += eps
            minus[pick] -= eps
            args_p = [plus if a is arr else a for a in (lv, dw, cw, cb)]
            args_m = [minus if a is arr else a for a in (lv, dw, cw, cb)]
            fd = (objective(*args_p) - objective(*args_m)) / (float(plus[pick]) - float(minus[pick]))
            assert abs(fd - float(grad[pick])) <= 1e-3 * max(abs(fd), 1.0), (pick, fd, float(grad[pick]))


def test_cross_level_tables():
    pos = cases.box_surface(1024, 7)
    fine = lo.build_lattice(pos, [0.05] * 3)
    coarse = lo.build_lattice(pos, [0.1] * 3)
    up = lo.neighbour_table(coarse["keys"], fine["keys"], 1, 1)        # coarse query <- fine neighbours
    down = lo.neighbour_table(fine["keys"], coarse["keys"], -1, 1)     # fine query <- coarse neighbours
    # every link is seen from both sides in opposite slots
    for c in range(coarse["nv"]):
        for s in range(8):
            f = up[c, s]
            if f >= 0:
                assert down[f, s ^ 1] == c
        if up[c, 8] >= 0:
            assert down[up[c, 8], 8] == c
    even = np.all(np.concatenate([fine["keys"], -fine["keys"].sum(1, keepdims=True)], 1) % 2 == 0, axis=1)
    assert np.all(down[even, :8] == -2) and np.all(down[~even, 8] == -2)


def test_scale_constants_are_ptxas_folded_not_hardware_rsqrt():
    """The reference binary's elevate() scale factors are constants: ptxas folds rsqrt.approx.ftz with
    the correctly rounded 1/sqrt (seen in the SASS), which differs from the MUFU.RSQ hardware result
    measured on the B200 (rsqrt_approx.json) for pos_dim 5."""
    import ctypes
    import struct
    def f32(bits):
        return np.frombuffer(struct.pack("<I", bits), dtype=np.float32)[0]
    for d, c_bits in ((3, 0x405105EC), (5, 0x409CC471)):
        got = (ctypes.c_uint32 * d)()
        lo._c().oracle_get_scale_table(ctypes.c_int(d), got)
        exact = [np.float32(np.float32(1.0 / np.sqrt(np.float64((i + 1) * (i + 2)))) * f32(c_bits)).view(np.uint32) for i in range(d)]
        assert [int(x) for x in got] == [int(x) for x in exact]
    path = os.path.join(GOLD, "rsqrt_approx.json")
    if os.path.isfile(path):
        with open(path) as f:
            hw = [int(b, 16) for b in json.load(f)["bits"]]
        hw_scale5 = [int(np.float32(f32(b) * f32(0x409CC471)).view(np.uint32)) for b in hw]
        got5 = (ctypes.c_uint32 * 5)()
        lo._c().oracle_get_scale_table(ctypes.c_int(5), got5)
        assert hw_scale5 != [int(x) for x in got5]          # the hardware values would give different keys


@pytest.mark.parametrize("name", list(cases.CASES))
def test_oracle_reproduces_reference_kernels(name):
    """PINS the oracle: bit-exact integers, bit-equal weights, values within fp32 tolerance."""
    g = _gold(name)
    pos, sig = g["positions"], g["sigmas"]
    n, d = pos.shape
    L = lo.build_lattice(pos, sig)
    assert L["nv"] == int(g["nv"])
    assert np.array_equal(L["keys"], g["keys"])
    assert np.array_equal(L["indices"], g["indices"])
    assert bits_equal(L["weights"], g["weights"]) == 0
    assert_close(lo.splat_accumulate(g["splat_in"], L["indices"], L["weights"], L["nv"]), g["splat_values"], 1e-5, "splat")
    assert bits_equal(lo.distribute_rows(pos, sig, g["distribute_in"], g["distribute_weights"]), g["distributed"]) == 0
    assert np.array_equal(g["distribute_indices"], L["indices"])
    for V in (1, 8, 32):
        assert_close(lo.slice_fwd(g[f"lv{V}"], L["indices"], L["weights"], n), g[f"slice{V}"], 1e-6, f"slice{V}")
        assert_close(lo.slice_bwd(g[f"slice_bwd_in{V}"], L["indices"], L["weights"], L["nv"]), g[f"slice_bwd{V}"], 1e-5, f"slice_bwd{V}")
    assert np.array_equal(g["slice_nop_indices"], L["indices"])
    assert_close(lo.gather_fwd(g["lv8"], L["indices"], L["weights"], n), g["gather8"], 1e-6, "gather")
    assert_close(lo.gather_bwd(g["gather_bwd_in8"], L["indices"], L["weights"], L["nv"], 8), g["gather_bwd8"], 1e-5, "gather_bwd")
    if "sc_logits" in g:
        logits, _ = lo.slice_classify_fwd(g["lv32"], L["indices"], L["weights"], g["sc_dw"], g["sc_w"], g["sc_b"], n)
        assert_close(logits, g["sc_logits"], 1e-5, "slice_classify")
        e = lo.slice_classify_bwd(g["sc_grad_in"], g["lv32"], L["indices"], L["weights"], g["sc_dw"], g["sc_w"], n)
        for a, key in zip(e, ("sc_g_lv", "sc_g_dw", "sc_g_w", "sc_g_b")):
            assert_close(a, g[key], 1e-4, key)
    F = 2 * (d + 1) + 1

    def rowidx(table, flip, id0):
        # golden stores the reference's raw 0 cells (vertex id 0 or never written) as -3
        r = lo.im2rowindices(table, 1, flip).reshape(-1, F).copy()
        src = table if not flip else np.concatenate([table[:, :F - 1].reshape(-1, (F - 1) // 2, 2)[:, :, ::-1].reshape(-1, F - 1), table[:, F - 1:]], 1)
        r[(src == -2) | (src == id0)] = -3
        return r

    for dil in (1, 2):
        T = lo.neighbour_table(L["keys"], L["keys"], 0, dil)
        for flip in (0, 1):
            assert np.array_equal(rowidx(T, bool(flip), int(g["id0_fine"])), g[f"rowidx_d{dil}_f{flip}"])
    T = lo.neighbour_table(L["keys"], L["keys"], 0, 1)
    assert_close(lo.conv_fwd(g["lv8"], T, g["conv_filter"]), g["conv8_16"], 1e-5, "conv")
    assert_close(lo.row2im(lo.im2row(g["lv8"], T), T, 8), g["row2im8"], 1e-6, "row2im")
    C = lo.build_lattice(pos, sig * 2.0)
    assert C["nv"] == int(g["coarse_nv"]) and np.array_equal(C["keys"], g["coarse_keys"])
    up = lo.neighbour_table(C["keys"], L["keys"], 1, 1)
    down = lo.neighbour_table(L["keys"], C["keys"], -1, 1)
    assert np.array_equal(rowidx(up, False, int(g["id0_fine"])), g["rowidx_coarse_from_fine"])
    assert np.array_equal(rowidx(down, False, int(g["id0_coarse"])), g["rowidx_fine_from_coarse"])
    assert np.array_equal(rowidx(down, True, int(g["id0_coarse"])), g["rowidx_fine_from_coarse_flip"])
    assert_close(lo.conv_fwd(g["lv8"], up, g["conv_filter"]), g["coarsen_conv8_16"], 1e-5, "coarsen conv")
    assert np.array_equal(lo.coarsen_keys(L["keys"]), g["coarsen_kernel_keys"])
