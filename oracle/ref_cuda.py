"""TEST INFRASTRUCTURE (GPU oracle) -- runs the reference's OWN device kernels on the B200.

oracle/build_ref.py compiles the unmodified /root/reference/include/lattice_net/kernels/LatticeGPU.cuh
with NVRTC (`-std=c++11 --use_fast_math`, the reference's jitify options) into oracle/_ref/lattice_ref.ptx.
This module loads that PTX through the CUDA driver (the same cuModuleLoadData path jitify uses,
/root/reference/deps/jitify/jitify.hpp:1001-1003) and launches the kernels with the reference's grids
(256-thread blocks, one thread per point / vertex, LatticeGPU.cuh:42-412) after the reference's host
preparation (zero / -1 fills, positions/sigma, `mm`; /root/reference/src/Lattice.cu).

Only tests/, __graft_entry__.smoke(), oracle/make_golden.py and bench.py's reference leg import this.
Needs a GPU; /root/reference is NOT needed at run time.
"""
import ctypes
import json
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
BLOCK = 256


class HashTableGPU(ctypes.Structure):
    """By-value kernel argument, layout of HashTableGPU.cuh:23-28 (48 bytes)."""
    _fields_ = [("m_capacity", ctypes.c_int), ("m_keys", ctypes.c_void_p), ("m_values", ctypes.c_void_p),
                ("m_entries", ctypes.c_void_p), ("m_nr_filled", ctypes.c_void_p), ("m_pos_dim", ctypes.c_int)]


def available():
    return os.path.isfile(os.path.join(REF_DIR, "lattice_ref.ptx")) and os.path.isfile(os.path.join(REF_DIR, "lattice_ref.names.json"))


class RefKernels:
    _instance = None

    @classmethod
    def get(cls):
        if cls._instance is None:
            cls._instance = RefKernels()
        return cls._instance

    def __init__(self):
        from cuda.bindings import driver
        self.drv = driver
        if not available():
            raise RuntimeError("oracle/_ref/lattice_ref.ptx missing: run `python oracle/build_ref.py` where /root/reference exists")
        torch.zeros(1, device="cuda")   # make torch create / bind the primary context
        with open(os.path.join(REF_DIR, "lattice_ref.names.json")) as f:
            self.meta = json.load(f)
        with open(os.path.join(REF_DIR, "lattice_ref.ptx"), "rb") as f:
            ptx = f.read() + b"\0"
        err, self.module = driver.cuModuleLoadData(ptx)   # driver JIT: PTX (.target sm_100) -> SASS
        self._chk(err, "cuModuleLoadData")
        self._fn = {}

    def _chk(self, err, what):
        if int(err) != 0:
            raise RuntimeError(f"{what} failed with CUresult {int(err)}")

    def has(self, name):
        return name in self.meta["names"]

    def function(self, name):
        if name not in self._fn:
            if name not in self.meta["names"]:
                raise KeyError(f"reference kernel instantiation {name} was not built (see oracle/build_ref.py)")
            err, fn = self.drv.cuModuleGetFunction(self.module, self.meta["names"][name].encode())
            self._chk(err, f"cuModuleGetFunction({name})")
            self._fn[name] = fn
        return self._fn[name]

    def launch(self, name, n_threads, args, smem=0):
        """args: list of ctypes objects (c_void_p, c_int, c_bool, HashTableGPU ...)."""
        fn = self.function(name)
        grid = max((n_threads - 1) // BLOCK + 1, 1)
        holders = list(args)
        ptrs = (ctypes.c_void_p * len(holders))(*[ctypes.addressof(a) for a in holders])
        stream = torch.cuda.current_stream().cuda_stream
        (err,) = self.drv.cuLaunchKernel(fn, grid, 1, 1, BLOCK, 1, 1, smem, stream, ctypes.addressof(ptrs), 0)
        self._chk(err, f"cuLaunchKernel({name})")


def _p(t):
    return ctypes.c_void_p(t.data_ptr() if t is not None else 0)


class RefTable:
    """HashTable (HashTable.cu:21-57): four device tensors + clear()."""

    def __init__(self, capacity, pos_dim, val_dim=1):
        dev = "cuda"
        self.capacity, self.pos_dim = capacity, pos_dim
        self.keys = torch.zeros((capacity, pos_dim), dtype=torch.int32, device=dev)
        self.values = torch.zeros((capacity, val_dim), dtype=torch.float32, device=dev)
        self.entries = torch.zeros((capacity,), dtype=torch.int32, device=dev)
        self.nr_filled = torch.zeros((1,), dtype=torch.int32, device=dev)
        self.clear()

    def clear(self):
        self.values.fill_(0)
        self.keys.fill_(0)
        self.entries.fill_(-1)
        self.nr_filled.fill_(0)

    def struct(self, values=None):
        v = self.values if values is None else values
        return HashTableGPU(self.capacity, self.keys.data_ptr(), v.data_ptr(), self.entries.data_ptr(),
                            self.nr_filled.data_ptr(), self.pos_dim)

    def nv(self):
        return int(self.nr_filled.item())


class RefLattice:
    """The reference's Lattice methods, restated as host prep + launches of the reference kernels."""

    def __init__(self, capacity, sigmas, lvl=1):
        self.k = RefKernels.get()
        self.capacity = capacity
        self.sigmas = torch.tensor(sigmas, dtype=torch.float32, device="cuda")
        self.lvl = lvl
        self.table = None
        self.values = None   # [nv x V] (or [C x V] right after a splat)

    # -- Lattice::splat_standalone / just_create_verts (Lattice.cu:196-290)
    def build(self, positions_raw, with_tables=True):
        n, d = positions_raw.shape
        if self.table is None:
            self.table = RefTable(self.capacity, d)
        idx = torch.empty((n * (d + 1),), dtype=torch.int32, device="cuda").fill_(-1)
        w = torch.empty((n * (d + 1),), dtype=torch.float32, device="cuda").fill_(-1)
        positions = positions_raw / self.sigmas
        self.k.launch(f"kernel_splat<{d},1>", n,
                      [_p(positions), ctypes.c_int(n), _p(idx), _p(w), self.table.struct(), ctypes.c_bool(with_tables)])
        return idx, w

    def splat(self, positions_raw, values):
        n, d = positions_raw.shape
        v = values.shape[1]
        self.table = RefTable(self.capacity, d, v)
        idx, w = self.build(positions_raw)
        self.k.launch(f"splatCacheNaive<{d},{v}>", n, [ctypes.c_int(n), _p(values), _p(idx), _p(w), self.table.struct()])
        self.values = self.table.values
        return idx, w

    # -- Lattice::distribute (Lattice.cu:351-410)
    def distribute(self, positions_raw, values):
        n, d = positions_raw.shape
        v = values.shape[1]
        self.table = RefTable(self.capacity, d, v)
        distributed = torch.zeros((n * (d + 1), d + v + 1), dtype=torch.float32, device="cuda")
        idx = torch.empty((n * (d + 1),), dtype=torch.int32, device="cuda").fill_(-1)
        w = torch.empty((n * (d + 1),), dtype=torch.float32, device="cuda").fill_(-1)
        positions = positions_raw / self.sigmas
        self.k.launch(f"distribute<{d},{v}>", n,
                      [_p(positions), _p(values), ctypes.c_int(n), _p(idx), _p(w), _p(distributed), self.table.struct()])
        return distributed, idx, w

    def nv(self):
        return self.table.nv()

    # -- Lattice::create_coarse_verts (Lattice.cu:670-703)
    def create_coarse_verts(self):
        d = self.table.pos_dim
        coarse = RefLattice(self.capacity, (self.sigmas * 2.0).tolist(), self.lvl + 1)
        coarse.table = RefTable(self.capacity, d)
        self.k.launch(f"coarsen<{d}>", self.capacity, [ctypes.c_int(self.capacity), self.table.struct(), coarse.table.struct()])
        return coarse

    # -- Lattice::create_coarse_verts_naive (Lattice.cu:706-740)
    def create_coarse_verts_naive(self, positions_raw):
        coarse = RefLattice(self.capacity, (self.sigmas * 2.0).tolist(), self.lvl + 1)
        coarse.build(positions_raw, with_tables=False)
        return coarse

    # -- Lattice::im2row (Lattice.cu:612-644)
    def im2row(self, neighbours, values_n, dilation=1, flip=False):
        nv = self.nv()
        d = self.table.pos_dim
        v = values_n.shape[1]
        F = 2 * (d + 1) + 1
        rowified = torch.zeros((nv, F * v), dtype=torch.float32, device="cuda")
        self.k.launch(f"im2row<{d},{v}>", nv,
                      [ctypes.c_int(nv), _p(rowified), ctypes.c_int(F), ctypes.c_int(dilation), self.table.struct(),
                       neighbours.table.struct(values_n), ctypes.c_int(self.lvl), ctypes.c_int(neighbours.lvl),
                       ctypes.c_bool(flip), ctypes.c_bool(False)])
        return rowified

    def im2rowindices(self, neighbours, val_dim, dilation=1, flip=False):
        nv = self.nv()
        d = self.table.pos_dim
        F = 2 * (d + 1) + 1
        rowified = torch.zeros((nv, F * val_dim), dtype=torch.int32, device="cuda")
        self.k.launch(f"im2rowindices<{d},{val_dim}>", nv,
                      [ctypes.c_int(nv), _p(rowified), ctypes.c_int(F), ctypes.c_int(dilation), self.table.struct(),
                       neighbours.table.struct(), ctypes.c_int(self.lvl), ctypes.c_int(neighbours.lvl),
                       ctypes.c_bool(flip), ctypes.c_bool(False)])
        return rowified

    # -- Lattice::convolve_im2row_standalone (Lattice.cu:424-474): im2row + fp32 mm
    def convolve(self, filter_bank, neighbours, values_n, dilation=1, flip=False):
        return self.im2row(neighbours, values_n, dilation, flip).mm(filter_bank)

    # -- Lattice::row2im (Lattice.cu:646-667)
    def row2im(self, rowified, neighbours, val_dim, dilation=1):
        nv = self.nv()
        d = self.table.pos_dim
        F = 2 * (d + 1) + 1
        out = torch.zeros((nv, val_dim), dtype=torch.float32, device="cuda")
        self.k.launch(f"row2im<{d},{val_dim}>", self.capacity,
                      [ctypes.c_int(self.capacity), _p(rowified), ctypes.c_int(F), ctypes.c_int(dilation),
                       self.table.struct(out), neighbours.table.struct(), ctypes.c_int(self.lvl),
                       ctypes.c_int(neighbours.lvl), ctypes.c_bool(False)])
        return out

    # -- slice family (Lattice.cu:744-1142)
    def slice_with_precomputation(self, positions_raw, values, idx, w):
        n, d = positions_raw.shape
        v = values.shape[1]
        out = torch.zeros((n, v), dtype=torch.float32, device="cuda")
        positions = positions_raw / self.sigmas
        self.k.launch(f"slice_with_precomputation<{d},{v}>", n,
                      [_p(positions), _p(out), ctypes.c_int(n), _p(idx), _p(w), self.table.struct(values)])
        return out

    def slice_no_precomputation(self, positions_raw, values):
        n, d = positions_raw.shape
        v = values.shape[1]
        out = torch.zeros((n, v), dtype=torch.float32, device="cuda")
        idx = torch.empty((n * (d + 1),), dtype=torch.int32, device="cuda").fill_(-1)
        w = torch.empty((n * (d + 1),), dtype=torch.float32, device="cuda").fill_(-1)
        positions = positions_raw / self.sigmas
        self.k.launch(f"slice_no_precomputation<{d},{v}>", n,
                      [_p(positions), _p(out), ctypes.c_int(n), _p(idx), _p(w), self.table.struct(values)])
        return out, idx, w

    def gather_with_precomputation(self, positions_raw, values, idx, w):
        n, d = positions_raw.shape
        v = values.shape[1]
        out = torch.zeros((n, (d + 1) * (v + 1)), dtype=torch.float32, device="cuda")
        positions = positions_raw / self.sigmas
        self.k.launch(f"gather_with_precomputation<{d},{v}>", n,
                      [_p(positions), _p(out), ctypes.c_int(n), _p(idx), _p(w), self.table.struct(values)])
        return out

    def slice_classify_with_precomputation(self, positions_raw, values, delta_w, cls_w, cls_b, idx, w):
        n, d = positions_raw.shape
        v = values.shape[1]
        nc = cls_w.shape[0]
        out = torch.zeros((n, nc), dtype=torch.float32, device="cuda")
        positions = positions_raw / self.sigmas
        self.k.launch(f"slice_classify_with_precomputation<{d},{v},{nc}>", n,
                      [_p(positions), _p(out), _p(delta_w), _p(cls_w), _p(cls_b), ctypes.c_int(n), _p(idx), _p(w),
                       self.table.struct(values)])
        return out

    def slice_backwards(self, grad_sliced, idx, w):
        n, v = grad_sliced.shape
        d = self.table.pos_dim
        grad = torch.zeros((self.nv(), v), dtype=torch.float32, device="cuda")
        self.k.launch(f"slice_backwards_with_precomputation_no_homogeneous<{d},{v}>", n,
                      [ctypes.c_int(n), _p(grad_sliced), _p(idx), _p(w), self.table.struct(grad)])
        return grad

    def gather_backwards(self, grad_gathered, idx, w):
        n = grad_gathered.shape[0]
        d = self.table.pos_dim
        v = grad_gathered.shape[1] // (d + 1) - 1
        grad = torch.zeros((self.nv(), v), dtype=torch.float32, device="cuda")
        self.k.launch(f"gather_backwards_with_precomputation<{d},{v}>", n,
                      [ctypes.c_int(n), _p(grad_gathered), _p(idx), _p(w), self.table.struct(grad)])
        return grad

    def slice_classify_backwards(self, grad_logits, values, delta_w, cls_w, cls_b, idx, w):
        n, nc = grad_logits.shape
        d = self.table.pos_dim
        v = values.shape[1]
        g_lv = torch.zeros_like(values)
        g_dw = torch.zeros_like(delta_w)
        g_w = torch.zeros_like(cls_w)
        g_b = torch.zeros_like(cls_b)
        self.k.launch(f"slice_classify_backwards_with_precomputation<{d},{v},{nc}>", n,
                      [ctypes.c_int(n), _p(grad_logits), _p(values), _p(idx), _p(w), _p(delta_w), _p(cls_w), _p(cls_b),
                       _p(g_lv), _p(g_dw), _p(g_w), _p(g_b), self.table.struct(values)])
        return g_lv, g_dw, g_w, g_b
