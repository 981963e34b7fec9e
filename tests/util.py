"""Helpers shared by the parity tests: canonical relabelling and tolerance checks."""
import numpy as np

from oracle import lattice_oracle as lo


def canonical(keys_np):
    """keys [nv,d] (GPU insertion order) -> (sorted UNIQUE keys, old_to_new, new_to_old).
    The reference's insert can, rarely, store the same key twice (its key compare reads through a
    non-coherent L1, HashTableGPU.cuh:457-463); duplicates map onto one canonical id and new_to_old
    names the first occurrence."""
    ks, first, inv = np.unique(np.asarray(keys_np), axis=0, return_index=True, return_inverse=True)
    return ks, inv.reshape(-1).astype(np.int64), first.astype(np.int64)


def agg(rows, old_to_new):
    """Per-vertex rows in GPU order -> canonical order, summing rows of duplicated vertices."""
    rows = np.asarray(rows)
    out = np.zeros((int(old_to_new.max()) + 1,) + rows.shape[1:], rows.dtype)
    np.add.at(out, old_to_new, rows)
    return out


def first(rows, new_to_old):
    """Per-vertex rows in GPU order -> canonical order for GATHER-type outputs (conv, im2row, row2im): every copy
    of a duplicated vertex holds the same row, so the first occurrence is taken instead of the sum."""
    return np.asarray(rows)[new_to_old]


def max_rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    return float(np.abs(a - b).max() / scale)


def assert_close(a, b, tol, what):
    err = max_rel_err(a, b)
    assert err <= tol, f"{what}: max rel err {err:.3e} > {tol:.1e}"


def bits_equal(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.uint32)
    b = np.ascontiguousarray(b, np.float32).view(np.uint32)
    return int((a != b).sum())
