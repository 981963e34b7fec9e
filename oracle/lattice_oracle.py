"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of AIS-Bonn/lattice_net's lattice operators.

numpy for the data movement, a small C file (oracle/lattice_oracle.c, compiled with gcc by
`build_c()`) for the bit-sensitive geometry.  Every function cites the reference lines it restates.
Vertex numbering: the reference numbers vertices in hash-insertion order, which is a race
(HashTableGPU.cuh:454); the oracle's canonical numbering is the lexicographic order of the keys,
and `canonical_order()` maps any GPU numbering onto it.

Parity status: PINNED by tests/golden/*.npz -- outputs of the reference's own kernels
(oracle/_ref, NVRTC build of the unmodified LatticeGPU.cuh) run on a B200 by oracle/make_golden.py.
"""
import ctypes
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_C_SRC = os.path.join(HERE, "lattice_oracle.c")
_C_LIB = os.path.join(HERE, "_build", "liboracle.so")
_RSQRT_JSON = os.path.join(HERE, "..", "tests", "golden", "rsqrt_approx.json")
_lib = None


def build_c(force=False):
    """gcc -O2 -ffp-contract=off: no FMA contraction, every fused op in the C file is an explicit fmaf."""
    if force or not os.path.isfile(_C_LIB) or os.path.getmtime(_C_LIB) < os.path.getmtime(_C_SRC):
        os.makedirs(os.path.dirname(_C_LIB), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", _C_SRC, "-o", _C_LIB, "-lm"])
    return _C_LIB


def _c():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build_c())
        lib.oracle_simplex.restype = ctypes.c_int
        lib.oracle_build_table.restype = ctypes.c_int
        lib.oracle_hash.restype = ctypes.c_uint32
        _lib = lib
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------------------------------
# geometry + structure
def simplex(positions_raw, sigmas):
    """kernel_splat geometry (LatticeGPU.cuh:718-806) after `positions_raw / sigmas` (Lattice.cu:226).
    Returns keys int32 [n, d+1, d], barycentric fp32 [n, d+1], scaled positions [n, d]."""
    pos = np.ascontiguousarray(positions_raw, dtype=np.float32)
    sig = np.ascontiguousarray(sigmas, dtype=np.float32)
    n, d = pos.shape
    keys = np.empty((n, d + 1, d), np.int32)
    bary = np.empty((n, d + 1), np.float32)
    scaled = np.empty((n, d), np.float32)
    rc = _c().oracle_simplex(_ptr(pos), _ptr(sig), ctypes.c_int(n), ctypes.c_int(d), _ptr(keys), _ptr(bary), _ptr(scaled))
    assert rc == 0
    return keys, bary, scaled


def lexsort_rows(keys):
    """Permutation that sorts key rows lexicographically (column 0 most significant)."""
    keys = np.asarray(keys)
    return np.lexsort(keys.T[::-1])


def build_lattice(positions_raw, sigmas):
    """kernel_splat / just_create_verts (LatticeGPU.cuh:707-842, Lattice.cu:196-290) with canonical
    vertex numbering.  Returns dict(keys [nv,d] sorted, indices [n*(d+1)], weights [n*(d+1)], nv)."""
    skeys, bary, scaled = simplex(positions_raw, sigmas)
    n, spv, d = skeys.shape
    flat = skeys.reshape(-1, d)
    uniq, inverse = np.unique(flat, axis=0, return_inverse=True)   # np.unique sorts rows lexicographically
    return dict(keys=uniq.astype(np.int32), indices=inverse.reshape(-1).astype(np.int32),
                weights=bary.reshape(-1).copy(), nv=int(uniq.shape[0]), scaled=scaled, simplex_keys=skeys)


def canonical_order(keys):
    """For keys [nv,d] in arbitrary (GPU insertion) order: returns (sorted_keys, old_to_new) so that
    sorted_keys[old_to_new[i]] == keys[i]."""
    keys = np.asarray(keys)
    order = lexsort_rows(keys)
    old_to_new = np.empty(len(order), np.int64)
    old_to_new[order] = np.arange(len(order))
    return keys[order], old_to_new


def relabel(indices, old_to_new):
    """Map a table of vertex ids (negative = absent, kept) through old_to_new."""
    idx = np.asarray(indices).astype(np.int64)
    out = idx.copy()
    m = idx >= 0
    out[m] = old_to_new[idx[m]]
    return out


def hash_chain_stats(keys, capacity):
    """Sequentially insert `keys` into the reference's table (HashTableGPU.cuh:35-54, 425-484) and
    report (nr_filled, longest probe chain); retrieve() gives up after 300 probes (:494)."""
    keys = np.ascontiguousarray(keys, np.int32)
    n, d = keys.shape
    tk = np.zeros((capacity, d), np.int32)
    te = np.empty(capacity, np.int32)
    chain = ctypes.c_int(0)
    nf = _c().oracle_build_table(_ptr(keys), ctypes.c_longlong(n), ctypes.c_int(d), ctypes.c_int(capacity), _ptr(tk), _ptr(te), None, ctypes.byref(chain))
    return nf, chain.value


def key_hash(keys):
    """HashTableGPU::hash (HashTableGPU.cuh:35-50), vectorised; uint32 wrap-around."""
    keys = np.asarray(keys, np.int64)
    k = np.zeros(keys.shape[0], np.uint64)
    for i in range(keys.shape[1]):
        k = (k + (keys[:, i] & 0xFFFFFFFF).astype(np.uint64)) & 0xFFFFFFFF
        k = (k * np.uint64(2531011)) & 0xFFFFFFFF
    return k.astype(np.uint32)


# --------------------------------------------------------------------------------------------------
# splat / slice family (all fp32 arithmetic, accumulation order = point order)
def splat_accumulate(values, indices, weights, nv):
    """splatCacheNaive (LatticeGPU.cuh:926-973): lattice[idx] += value * weight."""
    values = np.asarray(values, np.float32)
    n, v = values.shape
    spv = len(indices) // n
    out = np.zeros((nv, v), np.float32)
    idx = np.asarray(indices).reshape(n, spv)
    w = np.asarray(weights, np.float32).reshape(n, spv)
    for r in range(spv):
        m = idx[:, r] >= 0
        np.add.at(out, idx[m, r], values[m] * w[m, r][:, None])
    return out


def distribute_rows(positions_raw, sigmas, values, weights):
    """distribute (LatticeGPU.cuh:624-645): rows [scaled pos | value | barycentric]."""
    pos = np.asarray(positions_raw, np.float32) / np.asarray(sigmas, np.float32)
    values = np.asarray(values, np.float32)
    n, d = pos.shape
    spv = d + 1
    w = np.asarray(weights, np.float32).reshape(n, spv)
    out = np.zeros((n, spv, d + values.shape[1] + 1), np.float32)
    out[:, :, :d] = pos[:, None, :]
    out[:, :, d:d + values.shape[1]] = values[:, None, :]
    out[:, :, -1] = w
    return out.reshape(n * spv, -1)


def _fma32(a, b, c):
    # exact product in float64 (24+24 bits), one rounding to fp32 at the end up to double rounding
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def slice_fwd(lattice_values, indices, weights, n):
    """slice_with_precomputation (LatticeGPU.cuh:2552-2595): FMA chain over the simplex vertices."""
    lv = np.asarray(lattice_values, np.float32)
    spv = len(indices) // n
    idx = np.asarray(indices).reshape(n, spv)
    w = np.asarray(weights, np.float32).reshape(n, spv)
    out = np.zeros((n, lv.shape[1]), np.float32)
    for r in range(spv):
        m = idx[:, r] != -1
        out[m] = _fma32(lv[idx[m, r]], w[m, r][:, None], out[m])
    return out


def slice_bwd(grad_out, indices, weights, nv):
    """slice_backwards_with_precomputation_no_homogeneous (LatticeGPU.cuh:3540-3623)."""
    return splat_accumulate(grad_out, indices, weights, nv)


def gather_fwd(lattice_values, indices, weights, n):
    """gather_with_precomputation (LatticeGPU.cuh:2886-2929)."""
    lv = np.asarray(lattice_values, np.float32)
    v = lv.shape[1]
    spv = len(indices) // n
    idx = np.asarray(indices).reshape(n, spv)
    w = np.asarray(weights, np.float32).reshape(n, spv)
    out = np.zeros((n, spv, v + 1), np.float32)
    for r in range(spv):
        m = idx[:, r] >= 0
        out[m, r, :v] = lv[idx[m, r]] * w[m, r][:, None]
        out[m, r, v] = w[m, r]
    return out.reshape(n, spv * (v + 1))


def gather_bwd(grad_out, indices, weights, nv, val_dim):
    """gather_backwards_with_precomputation (LatticeGPU.cuh:3761-3817); the weight column's grad is dropped."""
    n = grad_out.shape[0]
    spv = len(indices) // n
    g = np.asarray(grad_out, np.float32).reshape(n, spv, val_dim + 1)
    idx = np.asarray(indices).reshape(n, spv)
    w = np.asarray(weights, np.float32).reshape(n, spv)
    out = np.zeros((nv, val_dim), np.float32)
    for r in range(spv):
        m = idx[:, r] >= 0
        np.add.at(out, idx[m, r], g[m, r, :val_dim] * w[m, r][:, None])
    return out


def slice_classify_fwd(lattice_values, indices, weights, delta_weights, cls_w, cls_b, n):
    """slice_classify_with_precomputation (LatticeGPU.cuh:3387-3464). float64 accumulation: this is
    the tolerance-checked (not bit-checked) part of the path."""
    lv = np.asarray(lattice_values, np.float64)
    spv = len(indices) // n
    idx = np.asarray(indices).reshape(n, spv)
    w = (np.asarray(weights, np.float32).reshape(n, spv) + np.asarray(delta_weights, np.float32).reshape(n, spv)).astype(np.float64)
    s = np.zeros((n, lv.shape[1]), np.float64)
    for r in range(spv):
        m = idx[:, r] >= 0
        s[m] += lv[idx[m, r]] * w[m, r][:, None]
    return (s @ np.asarray(cls_w, np.float64).T + np.asarray(cls_b, np.float64)).astype(np.float32), s


def slice_classify_bwd(grad_logits, lattice_values, indices, weights, delta_weights, cls_w, n):
    """slice_classify_backwards_with_precomputation (LatticeGPU.cuh:3628-3756), float64."""
    lv = np.asarray(lattice_values, np.float64)
    g = np.asarray(grad_logits, np.float64)
    W = np.asarray(cls_w, np.float64)
    spv = len(indices) // n
    idx = np.asarray(indices).reshape(n, spv)
    wd = (np.asarray(weights, np.float32).reshape(n, spv) + np.asarray(delta_weights, np.float32).reshape(n, spv)).astype(np.float64)
    t = g @ W                                   # [n, V]
    grad_lv = np.zeros_like(lv)
    grad_dw = np.zeros((n, spv), np.float64)
    s = np.zeros((n, lv.shape[1]), np.float64)
    for r in range(spv):
        m = idx[:, r] >= 0
        np.add.at(grad_lv, idx[m, r], t[m] * wd[m, r][:, None])
        grad_dw[m, r] = np.einsum("ij,ij->i", lv[idx[m, r]], t[m])
        s[m] += lv[idx[m, r]] * wd[m, r][:, None]
    grad_w = g.T @ s
    grad_b = g.sum(0)
    return grad_lv.astype(np.float32), grad_dw.astype(np.float32), grad_w.astype(np.float32), grad_b.astype(np.float32)


# --------------------------------------------------------------------------------------------------
# neighbourhood / convolution
def _key_lookup(keys):
    return {tuple(int(c) for c in k): i for i, k in enumerate(np.asarray(keys))}


def _round_half_away(x):
    return np.where(x >= 0, np.floor(x + 0.5), np.ceil(x - 0.5))


def neighbour_table(query_keys, nbr_keys, lvl_diff=0, dilation=1):
    """Traversal of im2row / im2rowindices (LatticeGPU.cuh:1479-1684).  Returns int64 [nv_q, F]:
    id, -1 (looked up, absent) or -2 (slot not examined by the reference)."""
    qk = np.asarray(query_keys, np.int64)
    nv, d = qk.shape
    F = 2 * (d + 1) + 1
    look = _key_lookup(nbr_keys)
    full = np.concatenate([qk, -qk.sum(1, keepdims=True)], 1).astype(np.float64)
    scale = float(2.0 ** lvl_diff)
    kf = full * scale
    out = np.full((nv, F), -2, np.int64)
    mm = scale if scale < 1.0 else 1.0
    for q in range(nv):
        all_int = True
        if scale < 1.0:
            all_int = bool(np.all(np.abs(kf[q] - np.trunc(kf[q])) <= 0.0001))
        if all_int:
            c = tuple(int(x) for x in _round_half_away(kf[q][:d]))
            idc = look.get(c, -1)
            if idc >= 0:
                out[q, F - 1] = idc
        if scale < 1.0 and all_int:
            continue
        for axis in range(d + 1):
            for sign_i, sgn in enumerate((1.0, -1.0)):
                nk = kf[q] + sgn * mm * dilation
                nk[axis] = kf[q][axis] - sgn * mm * dilation * d
                key = tuple(int(x) for x in _round_half_away(nk[:d]))
                out[q, 2 * axis + sign_i] = look.get(key, -1)
    return out


def im2row(nbr_values, table, flip=False):
    """im2row (LatticeGPU.cuh:1603-1684): [nv_q, F*V]; flip swaps the np/nm chunks."""
    vals = np.asarray(nbr_values, np.float32)
    nv, F = table.shape
    V = vals.shape[1]
    out = np.zeros((nv, F, V), np.float32)
    for c in range(F):
        src = (c ^ 1) if (flip and c < F - 1) else c
        ids = table[:, src]
        m = ids >= 0
        out[m, c] = vals[ids[m]]
    return out.reshape(nv, F * V)


def im2rowindices(table, val_dim, flip=False):
    """im2rowindices (LatticeGPU.cuh:1690-1920): ids replicated V times; untouched cells stay 0."""
    nv, F = table.shape
    out = np.zeros((nv, F, val_dim), np.int32)
    for c in range(F):
        src = (c ^ 1) if (flip and c < F - 1) else c
        ids = table[:, src].copy()
        ids[ids == -2] = 0
        out[:, c, :] = ids[:, None]
    return out.reshape(nv, F * val_dim)


def row2im(rowified, table, val_dim):
    """row2im (LatticeGPU.cuh:2196-2284): each vertex pulls the chunk its neighbours hold for it."""
    nv, F = table.shape
    rows = np.asarray(rowified, np.float32).reshape(-1, F, val_dim)
    out = np.zeros((nv, val_dim), np.float32)
    for slot in range(F):
        chunk = (slot ^ 1) if slot < F - 1 else slot
        ids = table[:, slot]
        m = ids >= 0
        out[m] += rows[ids[m], chunk]
    return out


def conv_fwd(nbr_values, table, filter_bank, flip=False):
    """convolve_im2row_standalone (Lattice.cu:424-474): im2row then fp32 GEMM (float64 accumulate here)."""
    rows = im2row(nbr_values, table, flip).astype(np.float64)
    return (rows @ np.asarray(filter_bank, np.float64)).astype(np.float32)


def conv_wgrad(nbr_values, table, grad_out):
    """lattice_funcs.py:302: grad_filter = im2row(values)^T . grad."""
    rows = im2row(nbr_values, table, False).astype(np.float64)
    return (rows.T @ np.asarray(grad_out, np.float64)).astype(np.float32)


def filter_for_dgrad(filter_bank, F, c_in, c_out):
    """lattice_funcs.py:304-311: [F*c_in, c_out] -> [F*c_out, c_in]."""
    fb = np.asarray(filter_bank).T.reshape(c_out, F, c_in).transpose(1, 0, 2)
    return np.ascontiguousarray(fb).reshape(F * c_out, c_in)


def coarsen_keys(fine_keys):
    """coarsen<d> (LatticeGPU.cuh:2314-2514): key set of the coarse lattice (sorted, unique)."""
    fk = np.asarray(fine_keys, np.int64)
    nv, d = fk.shape
    look = _key_lookup(fk)
    out = set()
    full = np.concatenate([fk, -fk.sum(1, keepdims=True)], 1)
    for v in range(nv):
        k = full[v]
        if np.any(k % 2 != 0):
            continue
        half = k // 2
        out.add(tuple(int(x) for x in half[:d]))
        for axis in range(d + 1):
            for sgn in (1, -1):
                nk = k + sgn
                nk[axis] = k[axis] - sgn * d
                if tuple(int(x) for x in nk[:d]) in look:
                    ck = half + sgn
                    ck[axis] = half[axis] - sgn * d
                    out.add(tuple(int(x) for x in ck[:d]))
    if not out:
        return np.zeros((0, d), np.int32)
    arr = np.array(sorted(out), np.int32)
    return arr
