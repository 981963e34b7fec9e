"""ctypes binding of liblattice_b200.so (include/lattice_b200.h).  There is no fallback: if the
CUDA library is missing or a call fails, this raises."""
import ctypes
import os

import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "liblattice_b200.so")

LN_OK = 0
LN_ERR_TABLE_FULL = -4
LN_ERR_VERTEX_BOUND = -5

_P = ctypes.c_void_p
_I = ctypes.c_int

# name -> argument ctypes (return type is int unless listed in _SPECIAL)
_SIGNATURES = {
    "ln_table_clear": [_P, _P, _P, _I, _P],
    "ln_table_status": [_P, _P, _P, _P, _P],
    "ln_splat_build": [_P, _P, _I, _I, _P, _P, _P, _P, _I, _I, _P, _P, _P],
    "ln_splat_accumulate": [_P, _P, _P, _I, _I, _I, _I, _P, _P],
    "ln_distribute": [_P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P],
    "ln_lookup_simplex": [_P, _P, _I, _I, _P, _P, _I, _I, _P, _P, _P],
    "ln_coarsen_keys": [_P, _P, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "ln_neighbour_table": [_P, _I, _P, _I, _P, _P, _I, _I, _I, _I, _P, _P],
    "ln_im2row": [_P, _P, _I, _I, _I, _I, _P, _P],
    "ln_im2rowindices": [_P, _I, _I, _I, _I, _P, _P],
    "ln_row2im": [_P, _P, _I, _I, _I, _P, _P],
    "ln_conv_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _I, _I, _P, _P],
    "ln_filter_prepare": [_P, _I, _I, _I, _I, _I, _P, _P],
    "ln_filter_prepare_batch": [_P, _I, ctypes.c_longlong, _P],
    "ln_conv_wgrad": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P],
    "ln_conv_bwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _I, _P, _I, _P, _I, _I, _I, _P],
    "ln_conv_bwd_join": [_P],
    "ln_filter_for_dgrad": [_P, _I, _I, _I, _P, _P],
    "ln_slice_fwd": [_P, _P, _P, _I, _I, _I, _I, _P, _P],
    "ln_slice_bwd": [_P, _P, _P, _I, _I, _I, _I, _P, _P],
    "ln_gather_fwd": [_P, _P, _P, _I, _I, _I, _P, _P],
    "ln_gather_bwd": [_P, _P, _P, _I, _I, _I, _P, _P],
    "ln_slice_classify_fwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P],
    "ln_slice_classify_bwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    "ln_scatter_max": [_P, _P, _I, _I, _I, _P, _P, _P, _P],
    "ln_scatter_sum_count": [_P, _P, _I, _I, _I, _P, _P, _P],
    "ln_group_norm_fwd": [_P, _P, _P, _I, _P, _I, _I, ctypes.c_float, _I, _P, _P, _P, _P],
    "ln_group_norm_bwd": [_P, _P, _P, _P, _P, _P, _I, _P, _I, _I, _I, _P, _P, _P, _P, _P],
}
_F = ctypes.c_float
_SIGNATURES.update({
    "ln_seg_loss_fwd": [_P, _P, _I, _I, _I, _P, _P, _P, _P],
    "ln_seg_loss_bwd": [_P, _P, _P, _P, _I, _I, _I, _P, _P],
    "ln_pointnet_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    "ln_pointnet_bwd": [_P, _P, _P, _P, _I, _I, _I, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    "ln_deltaw_fwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P],
    "ln_deltaw_bwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P],
    "ln_levels_status": [_P, _P, _I, _P, _P, _P],
    "ln_weight_norm_fwd": [_P, _P, _I, _I, _I, _P, _P],
    "ln_weight_norm_bwd": [_P, _P, _P, _I, _I, _I, _P, _P, _P],
    "ln_adamw_amsgrad": [_P, _P, _P, _P, _P, ctypes.c_longlong, _F, _F, _F, _F, _F, _F, _P, _P, _P],
})
_SPECIAL = {
    "ln_seg_loss_max_points": (ctypes.c_int, []),
    "ln_pointnet_supported": (ctypes.c_int, [_I, _I, _I, _I, _I]),
    "ln_pointnet_scratch_floats": (ctypes.c_longlong, [_I, _I, _I]),
    "ln_pointnet_grad_scratch_floats": (ctypes.c_longlong, [_I, _I, _I, _I, _I]),
    "ln_version": (ctypes.c_char_p, []),
    "ln_last_error": (ctypes.c_char_p, []),
    "ln_launch_count": (ctypes.c_longlong, []),
    "ln_conv_workspace_bytes": (ctypes.c_longlong, [_I, _I, _I, _I]),
    "ln_conv_needs_zero": (ctypes.c_int, [_I, _I, _I, _I, _I]),
    "ln_group_norm_workspace_bytes": (ctypes.c_longlong, [_I, _I, _I]),
    "ln_reset_launch_count": (None, []),
    "ln_set_programmatic_launch": (ctypes.c_int, [_I]),
}
EXPORTED_SYMBOLS = sorted(list(_SIGNATURES) + list(_SPECIAL))

_lib = None


class LatticeBackendError(RuntimeError):
    pass


def load():
    """Load the shared library (built by lattice_net_b200/build.py).  Fails loudly when absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise LatticeBackendError(
            f"{LIB_PATH} is missing: build it with `python -m lattice_net_b200.build` "
            "(there is no CPU or PyTorch fallback for the lattice kernels)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    for name, (restype, argtypes) in _SPECIAL.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not (t.is_cuda and t.is_contiguous()):
        raise LatticeBackendError("lattice kernels need contiguous CUDA tensors")
    return t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr(device=None):
    """cudaStream_t of torch's current stream on `device` (hot path: avoids building a Stream object)."""
    if _raw_stream is not None:
        if device is None:
            idx = torch.cuda.current_device()
        else:
            idx = device.index if isinstance(device, torch.device) else int(device)
            if idx is None:
                idx = torch.cuda.current_device()
        return _raw_stream(idx)
    return torch.cuda.current_stream(device).cuda_stream


_fn_cache = {}


def call(name, *args):
    """Call an int-returning entry point; raise with the library's message on failure."""
    fn = _fn_cache.get(name)
    if fn is None:
        fn = _fn_cache[name] = getattr(load(), name)
    rc = fn(*args)
    if rc != LN_OK:
        msg = load().ln_last_error().decode(errors="replace")
        raise LatticeBackendError(f"{name} failed (code {rc}): {msg}")
    return rc


def launch_count():
    return int(load().ln_launch_count())


def reset_launch_count():
    load().ln_reset_launch_count()


def version():
    return load().ln_version().decode()
