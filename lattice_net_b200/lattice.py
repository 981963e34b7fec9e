"""`Lattice` / `HashTable`: host-side mirror of the reference's pybind classes
(/root/reference/src/PyBridge.cxx:33-113, /root/reference/src/Lattice.cu, /root/reference/src/HashTable.cu).

Same method names, arguments, return tuples, dtypes and shapes as the reference; every device
operation goes through the C ABI of include/lattice_b200.h (hand-written sm_100a kernels).  PyTorch
is used only to own device memory and for the current stream.

What is deliberately different from the reference host code (results are identical):
  * positions are divided by sigma inside the kernels (no extra launch, Lattice.cu:226);
  * index / weight tables need no `fill_(-1)` pre-pass (Lattice.cu:214-215): every cell is written;
  * the vertex count is read back ONCE per lattice structure and shared by all handles that alias
    that structure (the reference re-syncs on every cloned handle, Lattice.cu:1326-1338 + :97-98);
  * the 1-hop neighbourhood is looked up once per (query, neighbour, dilation) and cached on the
    structure; convolutions consume the table directly and never materialise the im2row buffer
    (Lattice.cu:454-462);
  * the tensors' device is honoured (the reference hard-codes cuda:0), which scene-parallel
    multi-GPU training needs;
  * a full hash table raises instead of hanging (HashTableGPU.cuh:443-484 has no probe limit).
"""
import ctypes
import weakref

import torch

from . import _cabi
from ._cabi import call, ptr, stream_ptr
from .params import lattice_settings, parse_cfg

# conv arithmetic: 1 = tcgen05 3xTF32 error-compensated split (fp32-equivalent, the default), 2 = tcgen05 single-pass TF32,
# 0 = exact fp32 FMA on the CUDA cores (opt-in; also what layers with c_in % 32 != 0 run whatever the mode)
CONV_PRECISION = 1


def set_conv_precision(mode):
    global CONV_PRECISION
    assert mode in (0, 1, 2)
    CONV_PRECISION = mode


_SIGMA_CACHE = {}
_IDENTITY_TABLES = {}


# ---- prepared filters --------------------------------------------------------------------------------------------
# The tensor-core convolution reads a filter bank as PREPARED SLABS (pre-swizzled B tiles, TF32 high / low parts,
# csrc/ln_conv_tc.cu).  Weights change once per optimizer step, so the slabs of every bank -- its forward reading and
# its transposed reading for the data gradient -- are produced by ONE batched launch per step (`prepare_filters`)
# instead of by one kernel per convolution call.  A reading = (tensor, filter_extent, c_in, c_out, transposed) in the
# GEMM's own terms: K = filter_extent * c_in, N = c_out.
class _PreparedReading:
    __slots__ = ("slabs", "version", "precision", "dims")

    def __init__(self, slabs, dims):
        self.slabs, self.dims = slabs, dims
        self.version, self.precision = -1, -1


_PREPARED = {}          # (data_ptr, transposed) -> _PreparedReading
_JOB_TABLES = {}        # signature of a batch -> (device job table, n_jobs, total_threads)
_SCRATCH = {}           # (device index, stream) -> grow-only slab buffer for banks nobody prepared ahead of time


def _n_pad_sum(c_out):
    full, rest = divmod(int(c_out), 256)
    return full * 256 + (rest + 15) // 16 * 16


def _slab_floats(F, c_in, c_out):
    return 2 * F * c_in * _n_pad_sum(c_out)          # == ln_conv_workspace_bytes() / 4


def tensor_core_reading_ok(F, c_in, c_out):
    return CONV_PRECISION != 0 and c_in % 32 == 0 and 1 <= c_out <= 1024 and F >= 1     # conv_tc_supported(), ln_conv_tc.cu


def prepare_filters(readings):
    """readings: iterable of (tensor, filter_extent, c_in, c_out, transposed).  Prepares the tensor-core slabs of all of
    them in one launch (skipped entirely when every reading is current); shapes that do not run on the tensor cores are
    ignored.  Call it after the weights changed and before the convolutions that use them (LNN.forward and
    GraphedTrainStep do)."""
    import struct
    todo = []
    stale = False
    for t, F, c_in, c_out, transposed in readings:
        if not tensor_core_reading_ok(F, c_in, c_out) or not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            continue
        key = (t.data_ptr(), bool(transposed))
        dims = (int(F), int(c_in), int(c_out))
        entry = _PREPARED.get(key)
        if entry is None or entry.dims != dims or entry.slabs.device != t.device:
            entry = _PREPARED[key] = _PreparedReading(torch.empty((_slab_floats(*dims),), dtype=torch.float32, device=t.device), dims)
        if entry.version != t._version or entry.precision != CONV_PRECISION:
            stale = True
        todo.append((t, entry, key))
    if not todo:
        return 0
    device = todo[0][0].device
    capturing = torch.cuda.is_current_stream_capturing()
    sig = (CONV_PRECISION,) + tuple((k, e.dims, e.slabs.data_ptr()) for _, e, k in todo)
    table = _JOB_TABLES.get(sig)
    if table is None and capturing:
        raise RuntimeError("prepare_filters: this set of filter banks was never prepared outside a CUDA-graph capture (its job table "
                           "needs a host-to-device copy): run one warm-up pass of the same model call before capturing")
    if table is None:      # built on the first call with this set of banks, also when nothing is stale: a later capture needs it
        blob, first = b"", 0
        for t, e, (ptr_, transposed) in todo:
            F, c_in, c_out = e.dims
            blob += struct.pack("<QQiiiiiiq", t.data_ptr(), e.slabs.data_ptr(), F * c_in, c_in, c_out, 1 if transposed else 0,
                                1 if CONV_PRECISION == 1 else 0, 0, first)
            first += (F * c_in // 32) * ((_n_pad_sum(c_out) + 31) // 32)       # tiles of 32 x 32 slab elements, one CTA each
        host = torch.frombuffer(bytearray(blob), dtype=torch.uint8)
        if len(_JOB_TABLES) > 64:
            _JOB_TABLES.clear()
        table = _JOB_TABLES[sig] = (host.to(device), len(todo), first)
    if not (stale or capturing):
        return 0
    call("ln_filter_prepare_batch", ptr(table[0]), table[1], table[2], stream_ptr(device))
    for t, e, _ in todo:
        e.version, e.precision = t._version, CONV_PRECISION
    return len(todo)


def invalidate_prepared_filters():
    """Forget every prepared bank (needed only after writing to a weight through an alias that does not bump the
    tensor's version counter, e.g. `w.data`)."""
    _PREPARED.clear()
    _JOB_TABLES.clear()


def _slabs_for(filter_bank, F, c_in, c_out, transposed):
    """-> (slab tensor or None, prepared flag) for one convolution call."""
    if not tensor_core_reading_ok(F, c_in, c_out):
        return None, 0
    entry = _PREPARED.get((filter_bank.data_ptr(), bool(transposed)))
    if (entry is not None and entry.version == filter_bank._version and entry.precision == CONV_PRECISION
            and entry.dims == (F, c_in, c_out) and entry.slabs.device == filter_bank.device):
        return entry.slabs, 1
    device = filter_bank.device
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream_ptr(device))
    ws = _SCRATCH.get(key)
    need = _slab_floats(F, c_in, c_out)
    if ws is None or ws.numel() < need:
        ws = _SCRATCH[key] = torch.empty((max(need, 1 << 20),), dtype=torch.float32, device=device)
    return ws, 0


# ---- zeroed output arena --------------------------------------------------------------------------------------------
# On small lattices the convolution splits K across CTAs and combines the partial tiles with vector reductions, so its
# output must start at zero.  Inside a captured training step all such outputs are carved out of ONE buffer that a single
# memset clears at the start of the step (instead of one clearing launch per convolution).
class ZeroArena:
    def __init__(self, nr_floats, device):
        self.buf = torch.empty((max(int(nr_floats), 4),), dtype=torch.float32, device=device) if device is not None else None
        self.off = 0
        self.requested = 0          # floats asked for since the last reset (also counted when the arena is a dry run)

    def reset(self):
        if self.buf is not None:
            self.buf.zero_()
        self.off = self.requested = 0

    def take(self, rows, cols, device):
        n = rows * cols
        n_al = (n + 3) // 4 * 4      # 16-byte aligned rows for the vector reductions
        self.requested += n_al
        if self.buf is None or self.buf.device != device or self.off + n_al > self.buf.numel():
            return torch.zeros((rows, cols), dtype=torch.float32, device=device)
        out = self.buf[self.off:self.off + n].view(rows, cols)
        self.off += n_al
        return out


_ARENA = None


def set_zero_arena(arena):
    """Install (or with None remove) the arena `_zeroed` draws from; returns the previous one."""
    global _ARENA
    prev, _ARENA = _ARENA, arena
    return prev


def _zeroed(rows, cols, device):
    if _ARENA is not None:
        return _ARENA.take(rows, cols, device)
    return torch.zeros((rows, cols), dtype=torch.float32, device=device)


# ---- gradient targets --------------------------------------------------------------------------------------------
# parallel.GradBucket registers, per parameter, the slice of its flat gradient buffer; the backward passes of the lattice
# operators then write the weight gradients straight into those slices (cleared once per step by one memset) and hand
# autograd a fresh view of the slice, which it adopts as `.grad` without a copy.
_GRAD_TARGETS = {}      # data_ptr of the parameter -> (flat buffer, offset, shape)
_GRAD_TARGETS_ACTIVE = False


def register_grad_targets(targets):
    _GRAD_TARGETS.clear()
    _GRAD_TARGETS.update(targets)


def set_grad_targets_active(flag):
    global _GRAD_TARGETS_ACTIVE
    _GRAD_TARGETS_ACTIVE = bool(flag)


def grad_target(param):
    """A fresh view of the gradient-bucket slice of `param` (None when no zeroed bucket is active this step)."""
    if not _GRAD_TARGETS_ACTIVE:
        return None
    hit = _GRAD_TARGETS.get(param.data_ptr())
    if hit is None or tuple(hit[2]) != tuple(param.shape):
        return None
    flat, off, shape = hit
    return flat[off:off + param.numel()].view(shape)


def _conv_needs_zero(nv, F, c_in, c_out):
    return bool(_cabi.load().ln_conv_needs_zero(int(nv), int(F), int(c_in), int(c_out), int(CONV_PRECISION)))


def _check(cond, msg):
    # the reference aborts through loguru CHECK (Lattice.cu:162-181); here it is an exception
    if not cond:
        raise RuntimeError(msg)


def _check_dev_tensor(t, dtype, device, what):
    # the kernels reinterpret raw pointers: a wrong dtype / device must fail here, as data_ptr<int>() / data_ptr<float>()
    # do in the reference (Lattice.cu:751-787)
    if t.dtype != dtype:
        raise RuntimeError(f"{what} should be of type {dtype}, got {t.dtype}")
    if not t.is_cuda or (device is not None and t.device != device):
        raise RuntimeError(f"{what} should live on {device}, got {t.device}")


def _as_cuda_f32(t, device):
    if t.dtype != torch.float32:
        raise RuntimeError(f"expected a float32 tensor, got {t.dtype}")
    return t.to(device).contiguous()


def _conv_call(vals, table, fb, bias, residual, nv, F, c_in, c_out, flip, transposed):
    """One ln_conv_fwd call: picks up the prepared slabs of `fb` when there are any, a zeroed output where the kernel
    accumulates, and returns out [nv x c_out]."""
    dev = vals.device
    slabs, prepared = _slabs_for(fb, F, c_in, c_out, transposed)
    needs_zero = slabs is not None and _conv_needs_zero(nv, F, c_in, c_out)
    out = _zeroed(nv, c_out, dev) if needs_zero else torch.empty((nv, c_out), dtype=torch.float32, device=dev)
    if residual is not None:
        _check(tuple(residual.shape) == (nv, c_out) and residual.is_contiguous(), "residual should be a contiguous [nv x nr_filters] tensor")
        _check_dev_tensor(residual, torch.float32, dev, "residual")
    if bias is not None:
        _check_dev_tensor(bias, torch.float32, dev, "bias")
    call("ln_conv_fwd", ptr(vals), ptr(table), ptr(fb), ptr(bias), ptr(residual), nv, F, c_in, c_out, 1 if flip else 0,
         1 if transposed else 0, CONV_PRECISION, ptr(slabs), prepared, 1 if needs_zero else 0, ptr(out), stream_ptr(dev))
    return out


def _conv_bwd_call(nbr_vals, table_fwd, g, table_bwd, fb, nv_q, nv_n, F, c_in, c_out, fb_is_linear_weight, grad_param=None):
    """One ln_conv_bwd call -> (grad of the neighbour values or None, grad of the bank).  fb_is_linear_weight: `fb` is a
    torch.nn.Linear weight [c_out x c_in] (F = 1), i.e. the bank stored transposed.  grad_param: the parameter whose gradient
    the bank gradient IS (same layout); its gradient-bucket slice is written directly when one is active."""
    dev = g.device
    grad_in = None
    slabs, prepared, in_zero = None, 0, 0
    if table_bwd is not None:
        # the data gradient reads the bank transposed; for a Linear weight (stored transposed) that is the plain reading
        slabs, prepared = _slabs_for(fb, F, c_out, c_in, not fb_is_linear_weight)
        in_zero = 1 if (slabs is not None and _conv_needs_zero(nv_n, F, c_out, c_in)) else 0
        grad_in = _zeroed(nv_n, c_in, dev) if in_zero else torch.empty((nv_n, c_in), dtype=torch.float32, device=dev)
    target = grad_target(grad_param) if grad_param is not None else None
    if fb_is_linear_weight:
        _check(F == 1, "a Linear weight is a filter bank of extent 1")
        grad_filter = target if target is not None else torch.empty((c_out, c_in), dtype=torch.float32, device=dev)
    else:
        grad_filter = target if target is not None else torch.empty((F * c_in, c_out), dtype=torch.float32, device=dev)
    defer = 1 if (_DEFER_WGRAD_JOIN and grad_in is not None) else 0
    call("ln_conv_bwd", ptr(nbr_vals), ptr(table_fwd), ptr(g), ptr(table_bwd), ptr(fb), nv_q, nv_n, F, c_in, c_out, CONV_PRECISION,
         ptr(slabs), prepared, ptr(grad_in), in_zero, ptr(grad_filter), 1 if target is not None else 0, 1 if fb_is_linear_weight else 0, defer,
         stream_ptr(dev))
    if defer:
        _DEFERRED.append((nbr_vals, g, table_fwd))     # the side-stream kernel reads these: no reuse of their memory before the join
    return grad_in, grad_filter


# Deferred join of the weight-gradient stream (see ln_conv_bwd): off unless a caller that controls the whole backward pass
# (graphed.GraphedTrainStep) switches it on and calls join_deferred_wgrads() before the gradients are consumed.
_DEFER_WGRAD_JOIN = False
_DEFERRED = []


def set_defer_wgrad_join(flag):
    global _DEFER_WGRAD_JOIN
    prev, _DEFER_WGRAD_JOIN = _DEFER_WGRAD_JOIN, bool(flag)
    return prev


def join_deferred_wgrads(device):
    """Make the current stream of `device` wait for every weight gradient still running on the library's side stream."""
    if _DEFERRED:
        call("ln_conv_bwd_join", stream_ptr(device))
        _DEFERRED.clear()


def identity_table(nv, device):
    """[nv x 1] int32 table q -> q: with it the convolution kernels are a plain row-major GEMM (filter extent 1)."""
    key = (str(device), int(nv))
    t = _IDENTITY_TABLES.get(key)
    if t is None:
        if len(_IDENTITY_TABLES) > 256:
            _IDENTITY_TABLES.clear()
        t = _IDENTITY_TABLES[key] = torch.arange(nv, dtype=torch.int32, device=device).view(nv, 1)
    return t


def linear_forward(x, weight, bias=None, residual=None):
    """y = x W^T (+ bias) (+ residual) through the lattice-convolution kernels with filter extent 1 (tcgen05 when
    in_features % 32 == 0).  weight: torch.nn.Linear layout [out_features x in_features]."""
    m, k = x.shape
    n = int(weight.shape[0])
    return _conv_call(x.contiguous(), identity_table(m, x.device), weight, bias, residual, m, 1, k, n, False, True)


def linear_backward(x, weight, grad_out, need_input_grad=True, grad_param=None):
    """-> (dx = dy W or None, dW = dy^T x) with the data-gradient and weight-gradient kernels of the convolution."""
    m, k = x.shape
    n = int(weight.shape[0])
    table = identity_table(m, x.device)
    return _conv_bwd_call(x.contiguous(), table, grad_out.contiguous(), table if need_input_grad else None, weight, m, m, 1, k, n, True,
                          grad_param)


class _Structure:
    """Key set of one lattice level: the device arrays of HashTableGPU (HashTableGPU.cuh:23-28).
    Shared by every handle that aliases the same vertices (the reference shares the tensors between
    clones, Lattice.cu:89-95)."""

    def __init__(self, capacity, pos_dim, device, zero_keys=True, bound=None):
        self.capacity = int(capacity)
        self.pos_dim = int(pos_dim)
        self.device = device
        # static-shape mode: every per-vertex tensor of this level has exactly `bound` rows and the actual
        # vertex count stays on the device (nr_filled); nothing below ever syncs with the host
        self.bound = None if bound is None else min(int(bound), self.capacity)
        alloc = torch.zeros if zero_keys else torch.empty
        self.keys = alloc((self.capacity, self.pos_dim), dtype=torch.int32, device=device)
        self.entries = torch.empty((self.capacity,), dtype=torch.int32, device=device)
        self.nr_filled = torch.zeros((1,), dtype=torch.int32, device=device)
        self.status = torch.zeros((2,), dtype=torch.int32, device=device)
        self.nv = None            # host copy of nr_filled, None = dirty
        self.max_probe = 0
        self.neighbour_cache = {}
        self.version = 0          # bumped whenever the key set may have changed; neighbour tables INTO this structure record it
        self.clear()

    def clear(self):
        call("ln_table_clear", ptr(self.entries), ptr(self.nr_filled), ptr(self.status), self.capacity, stream_ptr(self.device))
        self.mark_dirty()

    def mark_dirty(self):
        self.nv = self.bound          # None = unknown until read back; a bound is known without asking the device
        self.neighbour_cache.clear()
        self.version += 1             # invalidates the tables OTHER structures hold into this one (checked on lookup)

    def nr_vertices_actual(self):
        """Blocking read of the device-side vertex count (raises on table overflow / exceeded bound)."""
        nv = ctypes.c_int(0)
        probe = ctypes.c_int(0)
        call("ln_table_status", ptr(self.nr_filled), ptr(self.status), ctypes.byref(nv), ctypes.byref(probe), stream_ptr(self.device))
        self.max_probe = int(probe.value)
        return int(nv.value)

    def nr_vertices(self):
        if self.nv is None:
            nv = ctypes.c_int(0)
            probe = ctypes.c_int(0)
            call("ln_table_status", ptr(self.nr_filled), ptr(self.status), ctypes.byref(nv), ctypes.byref(probe), stream_ptr(self.device))
            self.nv = int(nv.value)
            self.max_probe = int(probe.value)
        return self.nv

    def copy(self):
        other = _Structure.__new__(_Structure)
        other.capacity, other.pos_dim, other.device, other.bound = self.capacity, self.pos_dim, self.device, self.bound
        other.keys = self.keys.clone()
        other.entries = self.entries.clone()
        other.nr_filled = self.nr_filled.clone()
        other.status = self.status.clone()
        other.nv, other.max_probe, other.neighbour_cache, other.version = self.nv, self.max_probe, {}, 0
        return other


class HashTable:
    """Python-visible part of the reference's HashTable (PyBridge.cxx:33-39): `m_keys_tensor`,
    `m_nr_filled_tensor`; plus the per-handle values tensor."""

    def __init__(self, capacity):
        self.m_capacity = int(capacity)
        self.structure = None
        self.m_values_tensor = None

    @property
    def m_keys_tensor(self):
        return None if self.structure is None else self.structure.keys

    @property
    def m_entries_tensor(self):
        return None if self.structure is None else self.structure.entries

    @property
    def m_nr_filled_tensor(self):
        return None if self.structure is None else self.structure.nr_filled

    def is_initialized(self):
        return self.structure is not None

    def init(self, pos_dim, val_dim, device):
        # HashTable::init (HashTable.cu:21-47)
        self.structure = _Structure(self.m_capacity, pos_dim, device)
        self.m_values_tensor = torch.zeros((self.m_capacity, val_dim), dtype=torch.float32, device=device)

    def capacity(self):
        return self.m_capacity

    def pos_dim(self):
        _check(self.structure is not None, "hash table is not initialised: splat or distribute something first")
        return self.structure.pos_dim

    def val_dim(self):
        _check(self.m_values_tensor is not None, "hash table has no values yet")
        return int(self.m_values_tensor.shape[1])

    def set_values(self, new_values):
        self.m_values_tensor = new_values.contiguous()   # HashTable.cu:112-115: kept by reference


class Lattice:
    """The user-visible lattice handle (reference: class Lattice, /root/reference/src/Lattice.cu)."""

    SUPPORTS_TRANSPOSED_FILTER = True     # convolve_im2row_standalone(..., transposed_filter=True)

    m_expected_position_dimensions = -1   # static in the reference too (Lattice.cu:44,143)

    # ---- construction ------------------------------------------------------------------------
    def __init__(self, capacity=None, sigmas=None, name="", _clone_of=None):
        self.m_name = name
        self.m_lvl = 1
        self.m_positions = None
        self.m_sigmas = []
        self.m_sigmas_val_and_extent = []
        self._sigmas_dev = {}
        self.m_vertex_bounds = None          # static-shape mode: {level: rows}; see set_vertex_bounds
        if _clone_of is not None:
            return
        _check(capacity is not None and sigmas is not None, "Lattice needs a capacity and sigmas (or use Lattice.create(cfg))")
        self.m_hash_table = HashTable(int(capacity))
        self.set_sigmas(list(sigmas))

    @staticmethod
    def create(config, name=""):
        """Lattice.create(cfg_path[, name]) (PyBridge.cxx:46-47; Lattice::init_params Lattice.cu:107-132).
        `config` may also be an already parsed dict, or another Lattice (clone, Lattice.cu:73-101)."""
        if isinstance(config, Lattice):
            return config.clone_lattice()
        cfg = config if isinstance(config, dict) else parse_cfg(config)
        capacity, sigmas = lattice_settings(cfg)
        return Lattice(capacity, sigmas, name)

    def clone_lattice(self):
        # Lattice::Lattice(Lattice* other), Lattice.cu:73-101: shares structure + values, keeps level/sigmas
        other = Lattice(_clone_of=self)
        other.m_lvl = self.m_lvl
        other.m_sigmas = list(self.m_sigmas)
        other.m_sigmas_val_and_extent = list(self.m_sigmas_val_and_extent)
        other.m_positions = self.m_positions
        other.m_vertex_bounds = self.m_vertex_bounds
        other.m_hash_table = HashTable(self.m_hash_table.m_capacity)
        other.m_hash_table.structure = self.m_hash_table.structure
        other.m_hash_table.m_values_tensor = self.m_hash_table.m_values_tensor
        return other

    def set_vertex_bounds(self, bounds):
        """Static-shape mode (extension of the reference API; used for CUDA-graph capture of a whole step).
        `bounds` = rows per lattice level, a list [lvl1, lvl2, ...] or {lvl: rows}; None switches it off.
        Lattices built from this handle (distribute / splat / coarse levels) then allocate exactly that many
        rows, `nr_lattice_vertices()` returns the bound without a device sync, rows past the actual vertex count
        carry zeros through the network, and a cloud that needs more vertices than the bound is flagged in the
        structure's status word (its surplus vertices are dropped) instead of re-allocating."""
        if bounds is None:
            self.m_vertex_bounds = None
        elif isinstance(bounds, dict):
            self.m_vertex_bounds = {int(k): int(v) for k, v in bounds.items()}
        else:
            self.m_vertex_bounds = {i + 1: int(b) for i, b in enumerate(bounds)}

    def _bound_for(self, lvl):
        if self.m_vertex_bounds is None:
            return None
        _check(lvl in self.m_vertex_bounds, f"static-shape mode: no vertex bound was given for lattice level {lvl}")
        return self.m_vertex_bounds[lvl]

    def nr_lattice_vertices_actual(self):
        """Device-side vertex count (blocking); equals nr_lattice_vertices() outside static-shape mode."""
        return self._structure().nr_vertices_actual()

    def set_sigmas(self, sigmas_list):
        # Lattice::set_sigmas, Lattice.cu:134-160
        self.m_sigmas_val_and_extent = [(float(s), int(n)) for s, n in sigmas_list]
        self.m_sigmas = [s for s, n in self.m_sigmas_val_and_extent for _ in range(n)]
        Lattice.m_expected_position_dimensions = len(self.m_sigmas)
        self._sigmas_dev = {}

    def set_sigma(self, sigma):
        _check(len(self.m_sigmas_val_and_extent) == 1, "set_sigma assumes a single sigma group")   # Lattice.cu:1379
        self.m_sigmas = [float(sigma)] * len(self.m_sigmas)
        self._sigmas_dev = {}

    def increase_sigmas(self, stepsize):
        self.m_sigmas = [s + float(stepsize) for s in self.m_sigmas]
        self._sigmas_dev = {}

    # ---- small getters -------------------------------------------------------------------------
    def name(self):
        return self.m_name

    def set_name(self, name):
        self.m_name = name

    def val_dim(self):
        return self.m_hash_table.val_dim()

    def pos_dim(self):
        return self.m_hash_table.pos_dim()

    def capacity(self):
        return self.m_hash_table.capacity()

    def lvl(self):
        return self.m_lvl

    def positions(self):
        return self.m_positions

    def set_positions(self, positions_raw):
        self.m_positions = positions_raw

    def hash_table(self):
        return self.m_hash_table

    def values(self):
        return self.m_hash_table.m_values_tensor

    def sigmas_tensor(self):
        return torch.tensor(self.m_sigmas, dtype=torch.float32)

    def nr_lattice_vertices(self):
        _check(self.m_hash_table.structure is not None, "lattice has no vertices yet")
        return self.m_hash_table.structure.nr_vertices()

    def get_filter_extent(self, neighborhood_size):
        _check(neighborhood_size == 1, "only a 1-hop neighbourhood is implemented")   # Lattice.cu:1349
        return 2 * (self.pos_dim() + 1) + 1

    @staticmethod
    def get_expected_filter_extent(neighborhood_size):
        _check(neighborhood_size == 1, "only a 1-hop neighbourhood is implemented")
        _check(Lattice.m_expected_position_dimensions > 0, "create a Lattice first: the expected position dimension comes from its sigmas")
        return 2 * (Lattice.m_expected_position_dimensions + 1) + 1

    def set_values(self, new_values):
        # Lattice::set_values, Lattice.cu:1394-1398 (kept by reference; row count must equal the vertex count)
        ht = self.m_hash_table
        ht.m_values_tensor = new_values if new_values.is_contiguous() else new_values.contiguous()
        st = ht.structure
        nv = st.nv if (st is not None and st.nv is not None) else self.nr_lattice_vertices()
        if new_values.shape[0] != nv:
            raise RuntimeError(f"set_values: {new_values.shape[0]} rows but the lattice has {nv} vertices")

    # ---- helpers ---------------------------------------------------------------------------------
    def _device_of(self, t):
        if t is not None and t.is_cuda:
            return t.device
        st = self.m_hash_table.structure
        return st.device if st is not None else torch.device("cuda", torch.cuda.current_device())

    def _sigmas_on(self, device):
        # process-wide cache: coarse handles are re-created every scan, and an H2D copy per level would both
        # cost a launch and be illegal inside a CUDA-graph capture
        key = (str(device), tuple(self.m_sigmas))
        t = _SIGMA_CACHE.get(key)
        if t is None:
            t = _SIGMA_CACHE[key] = torch.tensor(self.m_sigmas, dtype=torch.float32, device=device)
        return t

    def _check_positions(self, positions_raw):
        # Lattice::check_positions, Lattice.cu:162-170
        _check(positions_raw.dtype == torch.float32, "positions should be of type float")
        _check(positions_raw.dim() == 2, f"positions should have dim 2, got sizes {tuple(positions_raw.shape)}")
        _check(len(self.m_sigmas) == positions_raw.shape[1],
               f"one sigma per position dimension is needed: {len(self.m_sigmas)} sigmas, pos_dim {positions_raw.shape[1]}")
        _check(positions_raw.is_contiguous(), "positions_raw is not contiguous, call .contiguous() on it")
        _check(positions_raw.shape[1] == Lattice.m_expected_position_dimensions,
               f"pos dim {positions_raw.shape[1]} differs from the expected {Lattice.m_expected_position_dimensions}")

    def _check_values(self, values):
        _check(values.dtype == torch.float32, "values should be of type float")
        _check(values.dim() == 2, f"values should have dim 2, got sizes {tuple(values.shape)}")
        _check(values.is_contiguous(), "values is not contiguous, call .contiguous() on it")

    def _structure(self):
        st = self.m_hash_table.structure
        _check(st is not None, "lattice has no vertices yet")
        return st

    def _neighbour_table(self, lattice_neighbours, dilation):
        """int32 [nv_query x F] vertex ids of `lattice_neighbours` around each vertex of self."""
        q, n = self._structure(), lattice_neighbours._structure()
        _check(abs(self.m_lvl - lattice_neighbours.m_lvl) <= 1,
               f"query lvl {self.m_lvl} and neighbour lvl {lattice_neighbours.m_lvl} may differ by one at most")   # Lattice.cu:442
        key = (id(n), int(dilation), self.m_lvl - lattice_neighbours.m_lvl)
        hit = q.neighbour_cache.get(key)
        # weak reference: no cycles between lattice levels; the neighbour's version: it may have been cleared,
        # re-splatted or grown in place since the table was built (the reference re-walks the hash table every call)
        if hit is not None and hit[0]() is n and hit[2] == n.version:
            return hit[1]
        nv = q.nr_vertices()
        _check(nv != 0, "why does this lattice have zero vertices?")   # Lattice.cu:443
        n.nr_vertices()   # surfaces a table overflow of the neighbour lattice before it is read
        F = 2 * (q.pos_dim + 1) + 1
        table = torch.empty((nv, F), dtype=torch.int32, device=q.device)
        call("ln_neighbour_table", ptr(q.keys), nv, ptr(q.nr_filled) if q.bound is not None else None, q.pos_dim,
             ptr(n.keys), ptr(n.entries), n.capacity, n.bound or 0,
             self.m_lvl - lattice_neighbours.m_lvl, int(dilation), ptr(table), stream_ptr(q.device))
        q.neighbour_cache[key] = (weakref.ref(n), table, n.version)
        return table

    # ---- splat / distribute -----------------------------------------------------------------------
    def begin_splat(self, reset_hashmap=True):
        # Lattice::begin_splat, Lattice.cu:185-193 (HashTable::clear / clear_only_values)
        ht = self.m_hash_table
        if ht.is_initialized():
            if ht.m_values_tensor is not None:
                if ht.m_values_tensor.shape[0] != ht.m_capacity:
                    ht.m_values_tensor = torch.zeros((ht.m_capacity, ht.m_values_tensor.shape[1]), dtype=torch.float32, device=ht.structure.device)
                else:
                    ht.m_values_tensor.zero_()
            if reset_hashmap:
                ht.structure.clear()

    def splat_standalone(self, positions_raw, values):
        """-> (splatting_indices i32 [N(d+1)], splatting_weights f32 [N(d+1)]); values() is [capacity x V]."""
        _check(positions_raw.shape[0] == values.shape[0], "positions and values need the same number of rows")
        self._check_positions(positions_raw)
        self._check_values(values)
        device = self._device_of(positions_raw)
        n, d = positions_raw.shape
        v = values.shape[1]
        self.m_positions = positions_raw
        ht = self.m_hash_table
        if not ht.is_initialized():
            ht.init(d, v, device)
        st = ht.structure
        st.bound = self._bound_for(self.m_lvl)
        pos = _as_cuda_f32(positions_raw, st.device)
        val = _as_cuda_f32(values, st.device)
        if ht.m_values_tensor is None or tuple(ht.m_values_tensor.shape) != (ht.m_capacity, v):
            ht.m_values_tensor = torch.zeros((ht.m_capacity, v), dtype=torch.float32, device=st.device)
        idx = torch.empty((n * (d + 1),), dtype=torch.int32, device=st.device)
        w = torch.empty((n * (d + 1),), dtype=torch.float32, device=st.device)
        s = stream_ptr(st.device)
        if n > 0:
            call("ln_splat_build", ptr(pos), ptr(self._sigmas_on(st.device)), n, d, ptr(st.keys), ptr(st.entries),
                 ptr(st.nr_filled), ptr(st.status), st.capacity, st.bound or 0, ptr(idx), ptr(w), s)
            call("ln_splat_accumulate", ptr(val), ptr(idx), ptr(w), n, d, v, 0, ptr(ht.m_values_tensor), s)
        st.mark_dirty()
        return idx, w

    def just_create_verts(self, positions_raw, return_indices_and_weights):
        self._check_positions(positions_raw)
        device = self._device_of(positions_raw)
        n, d = positions_raw.shape
        ht = self.m_hash_table
        if not ht.is_initialized():
            ht.init(d, 1, device)
        st = ht.structure
        st.bound = self._bound_for(self.m_lvl)
        pos = _as_cuda_f32(positions_raw, st.device)
        idx = w = None
        if return_indices_and_weights:
            idx = torch.empty((n * (d + 1),), dtype=torch.int32, device=st.device)
            w = torch.empty((n * (d + 1),), dtype=torch.float32, device=st.device)
        if n > 0:
            call("ln_splat_build", ptr(pos), ptr(self._sigmas_on(st.device)), n, d, ptr(st.keys), ptr(st.entries),
                 ptr(st.nr_filled), ptr(st.status), st.capacity, st.bound or 0, ptr(idx), ptr(w), stream_ptr(st.device))
        st.mark_dirty()
        return idx, w

    def distribute(self, positions_raw, values, reset_hashmap=True):
        """-> (distributed_lattice, distributed [N(d+1) x (d+V+1)], indices, weights)  (Lattice.cu:351-410)."""
        _check(positions_raw.shape[0] == values.shape[0], "positions and values need the same number of rows")
        self._check_positions(positions_raw)
        self._check_values(values)
        device = self._device_of(positions_raw)
        n, d = positions_raw.shape
        v = values.shape[1]
        self.m_positions = positions_raw
        ht = self.m_hash_table
        if not ht.is_initialized():
            ht.init(d, v, device)
        parent = ht.structure
        dev = parent.device
        pos = _as_cuda_f32(positions_raw, dev)
        val = _as_cuda_f32(values, dev)
        new = self.clone_lattice()
        new.m_name = "distributed_lattice"
        # the reference clones the parent's table and clears it (Lattice.cu:377-390); a fresh table is the same thing
        bound = self._bound_for(self.m_lvl)
        if reset_hashmap:
            new.m_hash_table.structure = _Structure(parent.capacity, d, dev, zero_keys=bound is None, bound=bound)
        else:
            new.m_hash_table.structure = parent.copy()
            new.m_hash_table.structure.bound = bound
        if bound is not None:
            # static-shape mode: the [capacity x V] placeholder of the reference is never read by the network
            # (PointNet replaces the values); a one-row stand-in avoids a capacity-sized clear per scan
            new.m_hash_table.m_values_tensor = torch.zeros((1, v), dtype=torch.float32, device=dev)
        elif ht.m_values_tensor is not None and ht.m_values_tensor.shape[1] == v:
            new.m_hash_table.m_values_tensor = torch.zeros_like(ht.m_values_tensor)
        else:
            new.m_hash_table.m_values_tensor = torch.zeros((ht.m_capacity, v), dtype=torch.float32, device=dev)
        st = new.m_hash_table.structure
        distributed = torch.empty((n * (d + 1), d + v + 1), dtype=torch.float32, device=dev)
        idx = torch.empty((n * (d + 1),), dtype=torch.int32, device=dev)
        w = torch.empty((n * (d + 1),), dtype=torch.float32, device=dev)
        if n > 0:
            call("ln_distribute", ptr(pos), ptr(self._sigmas_on(dev)), ptr(val), n, d, v, ptr(st.keys), ptr(st.entries),
                 ptr(st.nr_filled), ptr(st.status), st.capacity, st.bound or 0, ptr(idx), ptr(w), ptr(distributed), stream_ptr(dev))
        st.mark_dirty()
        return new, distributed, idx, w

    def distribute_structure(self, positions_raw, values, reset_hashmap=True):
        """`distribute` without its [N(d+1) x (d+V+1)] rows (extension): -> (distributed_lattice, indices, weights).  The
        fused PointNet front end (lattice_modules.PointNetModule.forward_fused) evaluates the rows on the fly instead."""
        _check(positions_raw.shape[0] == values.shape[0], "positions and values need the same number of rows")
        self._check_positions(positions_raw)
        self._check_values(values)
        device = self._device_of(positions_raw)
        n, d = positions_raw.shape
        v = values.shape[1]
        self.m_positions = positions_raw
        ht = self.m_hash_table
        if not ht.is_initialized():
            ht.init(d, v, device)
        parent = ht.structure
        dev = parent.device
        pos = _as_cuda_f32(positions_raw, dev)
        new = self.clone_lattice()
        new.m_name = "distributed_lattice"
        bound = self._bound_for(self.m_lvl)
        if reset_hashmap:
            new.m_hash_table.structure = _Structure(parent.capacity, d, dev, zero_keys=bound is None, bound=bound)
        else:
            new.m_hash_table.structure = parent.copy()
            new.m_hash_table.structure.bound = bound
        new.m_hash_table.m_values_tensor = torch.zeros((1, v), dtype=torch.float32, device=dev)      # replaced by the PointNet output
        st = new.m_hash_table.structure
        idx = torch.empty((n * (d + 1),), dtype=torch.int32, device=dev)
        w = torch.empty((n * (d + 1),), dtype=torch.float32, device=dev)
        if n > 0:
            call("ln_splat_build", ptr(pos), ptr(self._sigmas_on(dev)), n, d, ptr(st.keys), ptr(st.entries),
                 ptr(st.nr_filled), ptr(st.status), st.capacity, st.bound or 0, ptr(idx), ptr(w), stream_ptr(dev))
        st.mark_dirty()
        return new, idx, w

    def expand(self, positions_raw, point_multiplier, noise_stddev, expand_values):
        # Lattice::expand, Lattice.cu:292-348
        self._check_positions(positions_raw)
        device = self._device_of(positions_raw)
        d = positions_raw.shape[1]
        ht = self.m_hash_table
        if not ht.is_initialized():
            ht.init(d, 1, device)
        st0 = ht.structure
        pos = _as_cuda_f32(positions_raw, st0.device)
        expanded_pos = pos.repeat(point_multiplier, 1)
        expanded_pos = expanded_pos + torch.randn_like(expanded_pos) * noise_stddev
        new = self.clone_lattice()
        new.m_name = "expanded_lattice"
        new.m_hash_table.structure = st0.copy()
        new.m_hash_table.m_values_tensor = torch.zeros((1, self.val_dim()), dtype=torch.float32, device=st0.device)
        new.just_create_verts(expanded_pos.contiguous(), False)
        if expand_values:
            diff = new.nr_lattice_vertices() - self.nr_lattice_vertices()
            _check(diff >= 0, "expanding a lattice can only add vertices")
            new.set_values(torch.nn.functional.pad(self.values(), (0, 0, 0, diff)))
        return new

    # ---- coarse levels ---------------------------------------------------------------------------
    def _new_coarse_handle(self):
        st = self._structure()
        coarse = self.clone_lattice()
        coarse.m_name = "coarse_lattice"
        coarse.m_lvl = self.m_lvl + 1
        coarse.m_sigmas = [s * 2.0 for s in self.m_sigmas]   # Lattice.cu:678-682
        coarse.m_sigmas_val_and_extent = [(s * 2.0, n) for s, n in self.m_sigmas_val_and_extent]
        coarse._sigmas_dev = {}
        coarse.m_hash_table = HashTable(st.capacity)
        bound = coarse._bound_for(coarse.m_lvl)
        coarse.m_hash_table.structure = _Structure(st.capacity, st.pos_dim, st.device, zero_keys=bound is None, bound=bound)
        coarse.m_hash_table.m_values_tensor = torch.zeros((1, self.val_dim()), dtype=torch.float32, device=st.device)
        return coarse

    def create_coarse_verts(self):
        # Lattice::create_coarse_verts, Lattice.cu:670-703 (coarsen<d> kernel)
        st = self._structure()
        nv = st.nr_vertices()
        coarse = self._new_coarse_handle()
        cst = coarse.m_hash_table.structure
        call("ln_coarsen_keys", ptr(st.keys), ptr(st.entries), ptr(st.nr_filled), st.capacity, ptr(cst.keys),
             ptr(cst.entries), ptr(cst.nr_filled), ptr(cst.status), cst.capacity, cst.bound or 0, st.pos_dim, nv,
             stream_ptr(st.device))
        cst.mark_dirty()
        coarse.m_hash_table.m_values_tensor = torch.zeros((coarse.nr_lattice_vertices(), self.val_dim()), dtype=torch.float32, device=st.device)
        return coarse

    def create_coarse_verts_naive(self, positions_raw):
        # Lattice::create_coarse_verts_naive, Lattice.cu:706-740: splat the raw points at 2*sigma
        self._check_positions(positions_raw)
        coarse = self._new_coarse_handle()
        coarse.just_create_verts(positions_raw, False)
        return coarse

    # ---- convolution -----------------------------------------------------------------------------
    def convolve_im2row_standalone(self, filter_bank, dilation, lattice_neighbours=None, flip_neighbours=False, bias=None,
                                   transposed_filter=False, residual=None):
        """values_new[nv_self x nr_filters] = im2row(lattice_neighbours) . filter_bank, as one implicit-GEMM
        kernel (Lattice.cu:424-474).  Returns a new handle sharing this lattice's structure.
        transposed_filter (extension): filter_bank is the forward bank [F*nr_filters x val_dim] of the convolution
        whose data gradient this call computes; it is read transposed in place (lattice_funcs.py:304-311).
        bias / residual (extensions): added in the kernel's epilogue."""
        nbrs = self if lattice_neighbours is None else lattice_neighbours
        _check(filter_bank is not None and filter_bank.dim() == 2, "filter bank should be 2-D: (filter_extent*val_dim) x nr_filters")
        st = self._structure()
        vn = nbrs.val_dim()
        if transposed_filter:
            F = self.get_filter_extent(1)
            nr_filters = int(filter_bank.shape[0]) // F     # input width of the forward conv = output width here
            _check(nr_filters * F == filter_bank.shape[0] and int(filter_bank.shape[1]) == vn,
                   f"transposed filter bank should be [{F}*c x {vn}], got {tuple(filter_bank.shape)}")
        else:
            nr_filters = int(filter_bank.shape[1])
            F = int(filter_bank.shape[0]) // vn
            _check(F == self.get_filter_extent(1) and F * vn == filter_bank.shape[0],
                   f"filter extent should be {self.get_filter_extent(1)} but the filter bank has {filter_bank.shape[0]} rows for val_dim {vn}")
        table = self._neighbour_table(nbrs, dilation)
        nv = st.nr_vertices()
        vals = nbrs.values()
        _check(vals.shape[0] >= nbrs.nr_lattice_vertices(), "neighbour lattice values have fewer rows than vertices")
        fb = _as_cuda_f32(filter_bank, st.device)
        out = _conv_call(vals.contiguous(), table, fb, bias, residual, nv, F, vn, nr_filters, flip_neighbours, transposed_filter)
        new = self.clone_lattice()
        new.m_name = "convolved_lattice"
        new.m_hash_table.set_values(out)
        return new

    def conv_backward(self, lattice_neighbours, grad_values, filter_bank, dilation, need_input_grad=True, grad_param=None):
        """Backward of `out = self.convolve_im2row_standalone(filter_bank, dilation, lattice_neighbours)`:
        returns (grad w.r.t. lattice_neighbours.values(), grad w.r.t. filter_bank) from one C call
        (lattice_funcs.py:294-313 / 373-388 / 438-454 do it with two im2row buffers and three GEMMs)."""
        nbrs = self if lattice_neighbours is None else lattice_neighbours
        st, nst = self._structure(), nbrs._structure()
        table_fwd = self._neighbour_table(nbrs, dilation)
        nv_q, nv_n = st.nr_vertices(), nst.nr_vertices()
        vn = nbrs.val_dim()
        g = _as_cuda_f32(grad_values, st.device)
        _check(g.shape[0] == nv_q, "grad_values rows must match the query lattice")
        F = self.get_filter_extent(1)
        fb = _as_cuda_f32(filter_bank, st.device)
        _check(tuple(fb.shape) == (F * vn, int(g.shape[1])), f"filter bank should be [{F * vn} x {int(g.shape[1])}], got {tuple(fb.shape)}")
        table_bwd = nbrs._neighbour_table(self, dilation) if need_input_grad else None
        return _conv_bwd_call(nbrs.values().contiguous(), table_fwd, g, table_bwd, fb, nv_q, nv_n, F, vn, int(g.shape[1]), False, grad_param)

    def conv_weight_grad(self, lattice_neighbours, grad_values, filter_extent, dilation):
        """grad_filter = im2row(lattice_neighbours)^T . grad_values without the rowified buffer
        (lattice_funcs.py:298-302).  Extension of the reference API used by the autograd Functions."""
        nbrs = self if lattice_neighbours is None else lattice_neighbours
        st = self._structure()
        table = self._neighbour_table(nbrs, dilation)
        nv = st.nr_vertices()
        vn = nbrs.val_dim()
        g = _as_cuda_f32(grad_values, st.device)
        _check(g.shape[0] == nv, "grad_values rows must match the query lattice")
        grad_filter = torch.empty((filter_extent * vn, g.shape[1]), dtype=torch.float32, device=st.device)
        call("ln_conv_wgrad", ptr(nbrs.values().contiguous()), ptr(table), ptr(g), nv, filter_extent, vn,
             int(g.shape[1]), CONV_PRECISION, 0, ptr(grad_filter), stream_ptr(st.device))
        return grad_filter

    @staticmethod
    def filter_for_data_grad(filter_bank, filter_extent, val_dim):
        """[F*val_dim x nr_filters] -> [F*nr_filters x val_dim] (lattice_funcs.py:304-311) in one kernel."""
        fb = filter_bank.contiguous()
        nr_filters = int(fb.shape[1])
        out = torch.empty((filter_extent * nr_filters, val_dim), dtype=torch.float32, device=fb.device)
        call("ln_filter_for_dgrad", ptr(fb), filter_extent, val_dim, nr_filters, ptr(out), stream_ptr(fb.device))
        return out

    def im2row(self, lattice_neighbours, filter_extent, dilation, flip_neighbours):
        nbrs = self if lattice_neighbours is None else lattice_neighbours
        _check(filter_extent == self.get_filter_extent(1), f"filter extent should be {self.get_filter_extent(1)}")
        st = self._structure()
        table = self._neighbour_table(nbrs, dilation)
        nv, vn = st.nr_vertices(), nbrs.val_dim()
        out = torch.empty((nv, filter_extent * vn), dtype=torch.float32, device=st.device)
        call("ln_im2row", ptr(nbrs.values().contiguous()), ptr(table), nv, filter_extent, vn, 1 if flip_neighbours else 0,
             ptr(out), stream_ptr(st.device))
        return out

    def im2rowindices(self, lattice_neighbours, filter_extent, dilation, flip_neighbours):
        nbrs = self if lattice_neighbours is None else lattice_neighbours
        _check(filter_extent == self.get_filter_extent(1), f"filter extent should be {self.get_filter_extent(1)}")
        st = self._structure()
        table = self._neighbour_table(nbrs, dilation)
        nv, vn = st.nr_vertices(), nbrs.val_dim()
        out = torch.empty((nv, filter_extent * vn), dtype=torch.int32, device=st.device)
        call("ln_im2rowindices", ptr(table), nv, filter_extent, vn, 1 if flip_neighbours else 0, ptr(out), stream_ptr(st.device))
        return out

    def row2im(self, lattice_rowified, dilation, filter_extent, nr_filters, lattice_neighbours=None):
        # Lattice::row2im, Lattice.cu:646-667: replaces this handle's values
        nbrs = self if lattice_neighbours is None else lattice_neighbours
        _check(lattice_rowified.is_contiguous(), "lattice rowified is not contiguous, call .contiguous() on it")
        _check(lattice_rowified.shape[1] // filter_extent == self.val_dim(), "each row of the rowified lattice should be val_dim*filter_extent long")
        st = self._structure()
        table = self._neighbour_table(nbrs, dilation)
        nv, v = st.nr_vertices(), self.val_dim()
        out = torch.empty((nv, v), dtype=torch.float32, device=st.device)
        call("ln_row2im", ptr(lattice_rowified), ptr(table), nv, filter_extent, v, ptr(out), stream_ptr(st.device))
        self.m_hash_table.m_values_tensor = out
        return out

    # ---- slice family ------------------------------------------------------------------------------
    def _slice_prep(self, positions_raw, idx, w):
        self._check_positions(positions_raw)
        _check(self.val_dim() > 0, "splat something first so that there are values to slice")
        n, d = positions_raw.shape
        _check(d == self.pos_dim(), "position dimension differs from the one the lattice was created with")
        if idx is not None:
            _check(idx.numel() == n * (d + 1) and w.numel() == n * (d + 1),
                   f"indices / weights should have {n * (d + 1)} elements, got {tuple(idx.shape)} and {tuple(w.shape)}")   # Lattice.cu:771-773
            dev = self._structure().device
            _check_dev_tensor(idx, torch.int32, dev, "splatting_indices")
            _check_dev_tensor(w, torch.float32, dev, "splatting_weights")
            idx, w = idx.contiguous(), w.contiguous()
        return n, d, idx, w

    def slice_standalone_with_precomputation(self, positions_raw, splatting_indices, splatting_weights):
        n, d, idx, w = self._slice_prep(positions_raw, splatting_indices, splatting_weights)
        vals = self.values().contiguous()
        out = torch.empty((n, self.val_dim()), dtype=torch.float32, device=vals.device)
        call("ln_slice_fwd", ptr(vals), ptr(idx), ptr(w), n, d, self.val_dim(), int(vals.shape[0]), ptr(out), stream_ptr(vals.device))
        return out

    def slice_standalone_no_precomputation(self, positions_raw):
        n, d, _, _ = self._slice_prep(positions_raw, None, None)
        st = self._structure()
        pos = _as_cuda_f32(positions_raw, st.device)
        idx = torch.empty((n * (d + 1),), dtype=torch.int32, device=st.device)
        w = torch.empty((n * (d + 1),), dtype=torch.float32, device=st.device)
        call("ln_lookup_simplex", ptr(pos), ptr(self._sigmas_on(st.device)), n, d, ptr(st.keys), ptr(st.entries),
             st.capacity, st.bound or 0, ptr(idx), ptr(w), stream_ptr(st.device))
        return self.slice_standalone_with_precomputation(positions_raw, idx, w), idx, w

    def gather_standalone_with_precomputation(self, positions_raw, splatting_indices, splatting_weights):
        n, d, idx, w = self._slice_prep(positions_raw, splatting_indices, splatting_weights)
        vals = self.values().contiguous()
        v = self.val_dim()
        out = torch.empty((n, (d + 1) * (v + 1)), dtype=torch.float32, device=vals.device)
        call("ln_gather_fwd", ptr(vals), ptr(idx), ptr(w), n, d, v, ptr(out), stream_ptr(vals.device))
        return out

    def gather_standalone_no_precomputation(self, positions_raw):
        # the reference's kernel for this is ill-formed (LatticeGPU.cuh:2875); composing lookup + gather gives
        # what it evidently intends
        n, d, _, _ = self._slice_prep(positions_raw, None, None)
        st = self._structure()
        pos = _as_cuda_f32(positions_raw, st.device)
        idx = torch.empty((n * (d + 1),), dtype=torch.int32, device=st.device)
        w = torch.empty((n * (d + 1),), dtype=torch.float32, device=st.device)
        call("ln_lookup_simplex", ptr(pos), ptr(self._sigmas_on(st.device)), n, d, ptr(st.keys), ptr(st.entries),
             st.capacity, st.bound or 0, ptr(idx), ptr(w), stream_ptr(st.device))
        return self.gather_standalone_with_precomputation(positions_raw, idx, w), idx, w

    def slice_classify_with_precomputation(self, positions_raw, delta_weights, linear_clasify_weight, linear_clasify_bias,
                                           nr_classes, splatting_indices, splatting_weights):
        n, d, idx, w = self._slice_prep(positions_raw, splatting_indices, splatting_weights)
        vals = self.values().contiguous()
        dev = vals.device
        dw = _as_cuda_f32(delta_weights, dev)
        cw = _as_cuda_f32(linear_clasify_weight, dev)
        cb = _as_cuda_f32(linear_clasify_bias, dev)
        _check(tuple(cw.shape) == (nr_classes, self.val_dim()), "classifier weight should be nr_classes x val_dim")
        out = torch.empty((n, nr_classes), dtype=torch.float32, device=dev)
        call("ln_slice_classify_fwd", ptr(vals), ptr(idx), ptr(w), ptr(dw), ptr(cw), ptr(cb), n, d, self.val_dim(),
             int(nr_classes), ptr(out), stream_ptr(dev))
        return out

    def slice_classify_no_precomputation(self, positions_raw, delta_weights, linear_clasify_weight, linear_clasify_bias, nr_classes):
        n, d, _, _ = self._slice_prep(positions_raw, None, None)
        st = self._structure()
        pos = _as_cuda_f32(positions_raw, st.device)
        idx = torch.empty((n * (d + 1),), dtype=torch.int32, device=st.device)
        w = torch.empty((n * (d + 1),), dtype=torch.float32, device=st.device)
        call("ln_lookup_simplex", ptr(pos), ptr(self._sigmas_on(st.device)), n, d, ptr(st.keys), ptr(st.entries),
             st.capacity, st.bound or 0, ptr(idx), ptr(w), stream_ptr(st.device))
        logits = self.slice_classify_with_precomputation(positions_raw, delta_weights, linear_clasify_weight,
                                                         linear_clasify_bias, nr_classes, idx, w)
        return logits, idx, w

    def slice_backwards_standalone_with_precomputation(self, *args, **kwargs):
        # the reference's wrapper targets a kernel that is commented out (LatticeGPU.cuh:3467-3536)
        raise RuntimeError("slice_backwards_standalone_with_precomputation has no kernel in the reference either; "
                           "use slice_backwards_standalone_with_precomputation_no_homogeneous")

    def slice_backwards_standalone_with_precomputation_no_homogeneous(self, positions_raw, grad_sliced_values,
                                                                      splatting_indices, splatting_weights):
        # Lattice.cu:1067-1088: the handle's values become the gradient w.r.t. the lattice values
        n, d, idx, w = self._slice_prep(positions_raw, splatting_indices, splatting_weights)
        _check(grad_sliced_values.dim() == 2 and grad_sliced_values.is_contiguous(), "grad_sliced_values must be 2-D and contiguous")
        st = self._structure()
        _check_dev_tensor(grad_sliced_values, torch.float32, st.device, "grad_sliced_values")
        v = int(grad_sliced_values.shape[1])
        grad = _zeroed(st.nr_vertices(), v, st.device)
        call("ln_slice_bwd", ptr(grad_sliced_values), ptr(idx), ptr(w), n, d, v, int(grad.shape[0]), ptr(grad), stream_ptr(st.device))
        self.m_hash_table.m_values_tensor = grad

    def gather_backwards_standalone_with_precomputation(self, positions_raw, grad_sliced_values, splatting_indices, splatting_weights):
        # Lattice.cu:1117-1142
        n, d, idx, w = self._slice_prep(positions_raw, splatting_indices, splatting_weights)
        _check(grad_sliced_values.dim() == 2 and grad_sliced_values.is_contiguous(), "grad_sliced_values must be 2-D and contiguous")
        st = self._structure()
        _check_dev_tensor(grad_sliced_values, torch.float32, st.device, "grad_sliced_values")
        v = int(grad_sliced_values.shape[1]) // (d + 1) - 1
        grad = _zeroed(st.nr_vertices(), v, st.device)
        call("ln_gather_bwd", ptr(grad_sliced_values), ptr(idx), ptr(w), n, d, v, ptr(grad), stream_ptr(st.device))
        self.m_hash_table.m_values_tensor = grad

    def slice_classify_backwards_with_precomputation(self, grad_class_logits, positions_raw, initial_values, delta_weights,
                                                     linear_clasify_weight, linear_clasify_bias, nr_classes,
                                                     grad_lattice_values, grad_delta_weights, grad_linear_clasify_weight,
                                                     grad_linear_clasify_bias, splatting_indices, splatting_weights):
        # Lattice.cu:1091-1115: accumulates into the four caller-allocated (zeroed) gradients
        n, d, idx, w = self._slice_prep(positions_raw, splatting_indices, splatting_weights)
        _check(grad_class_logits.dim() == 2 and grad_class_logits.is_contiguous(), "grad_class_logits must be 2-D and contiguous")
        vals = initial_values.contiguous()
        dev = vals.device
        for t, what in ((grad_class_logits, "grad_class_logits"), (vals, "initial_values"), (delta_weights, "delta_weights"),
                        (linear_clasify_weight, "linear_clasify_weight")):
            _check_dev_tensor(t, torch.float32, dev, what)
        for t in (grad_lattice_values, grad_delta_weights, grad_linear_clasify_weight, grad_linear_clasify_bias):
            _check(t.is_contiguous(), "gradient buffers must be contiguous CUDA tensors")
            _check_dev_tensor(t, torch.float32, dev, "gradient buffers")
        call("ln_slice_classify_bwd", ptr(grad_class_logits), ptr(vals), ptr(idx), ptr(w), ptr(delta_weights.contiguous()),
             ptr(linear_clasify_weight.contiguous()), n, d, int(vals.shape[1]), int(nr_classes), ptr(grad_lattice_values),
             ptr(grad_delta_weights), ptr(grad_linear_clasify_weight), ptr(grad_linear_clasify_bias), stream_ptr(dev))
