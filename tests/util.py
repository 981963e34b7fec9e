"""Helpers shared by the parity tests: canonical relabelling and tolerance checks."""
import numpy as np

from oracle import lattice_oracle as lo


def canonical(keys_np):
    """keys [nv,d] (GPU insertion order) -> (sorted keys, old_to_new, new_to_old)."""
    ks, o2n = lo.canonical_order(keys_np)
    return ks, o2n, np.argsort(o2n)


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
