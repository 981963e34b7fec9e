"""TEST / BASELINE INFRASTRUCTURE -- the reference arm of bench.py (`--impl reference`).

The reference is CUDA-only (README.md:20) and its C++ host library cannot be built here (EasyPBR, Boost, Eigen, loguru,
configuru are absent, no network -- see DESIGN.md).  What runs UNMODIFIED:
  * its device code: oracle/_ref holds the NVRTC build of the reference's own LatticeGPU.cuh / HashTableGPU.cuh
    (oracle/build_ref.py), launched with the reference's grids (oracle/ref_cuda.py);
  * its Python layer: baseline/_ref/latticenet_py/{lattice/lattice_funcs, lattice_modules, lattice_wrapper, models,
    utils/utils}.py are byte-for-byte copies of the reference's files (oracle/install_ref_py.py) and are imported as they
    are -- the reference's autograd Functions, modules, LNN, weight-norm wrappers and initialisers.
What is restated here is only the piece in between, the compiled `latticenet` module (`RefHandle` below, following
/root/reference/src/Lattice.cu and src/HashTable.cu: C-sized `fill_` clears, the table clones of distribute(),
`positions / sigma` as a separate op, -1 fills of the index tables, a zeros im2row buffer + fp32 `torch.mm` per
convolution, one blocking D2H read of the vertex count per cloned handle, Lattice.cu:1326-1338), plus import shims for
modules that are not installed: `easypbr` (profiler no-ops), `torch_scatter` (torch scatter_reduce), `termcolor`.
None of this repo's kernels, modules or optimizer code is on that path.
"""
import ctypes
import json
import os
import time

import numpy as np
import torch

from .ref_cuda import RefKernels, RefTable, _p


class RefHandle:
    """Subset of the reference `Lattice` API needed by the lattice modules, on the reference kernels."""

    def __init__(self, capacity, sigmas, lvl=1):
        self.k = RefKernels.get()
        self.cap = capacity
        self.m_sigmas = list(sigmas)
        self.sigmas_t = torch.tensor(self.m_sigmas, dtype=torch.float32, device="cuda")
        self.m_lvl = lvl
        self.table = None
        self.m_values = None
        self.m_positions = None
        self.dirty = True          # HashTable::m_nr_filled_is_dirty starts true on every new handle
        self.nv_cached = -1

    # -- handle bookkeeping ----------------------------------------------------------------------
    def clone_lattice(self):
        o = RefHandle(self.cap, self.m_sigmas, self.m_lvl)
        o.table, o.m_values, o.m_positions = self.table, self.m_values, self.m_positions
        return o

    def nr_lattice_vertices(self):
        if self.dirty:                               # cudaMemcpy D2H, Lattice.cu:1333-1338
            self.nv_cached = int(self.table.nr_filled.item())
            self.dirty = False
        return self.nv_cached

    def set_values(self, v):
        self.m_values = v.contiguous()
        assert v.shape[0] == self.nr_lattice_vertices()

    def values(self):
        return self.m_values

    def val_dim(self):
        return int(self.m_values.shape[1])

    def pos_dim(self):
        return self.table.pos_dim

    def positions(self):
        return self.m_positions

    def lvl(self):
        return self.m_lvl

    def get_filter_extent(self, n):
        return 2 * (self.pos_dim() + 1) + 1

    m_expected_pos_dim = 3                       # static in the reference (Lattice.cu:44,143), set by the arm before the model is built

    @staticmethod
    def get_expected_filter_extent(n):
        return 2 * (RefHandle.m_expected_pos_dim + 1) + 1

    def set_val_dim(self, v):                    # bookkeeping only in the reference
        pass

    def name(self):
        return "lattice"

    def begin_splat(self, reset=True):
        if self.table is not None:                   # HashTable::clear: 4 fills over C-sized tensors
            self.table.clear()
            self.dirty = True

    def _struct(self, values=None):
        return self.table.struct(self.m_values if values is None else values)

    # -- distribute (Lattice.cu:351-410) ---------------------------------------------------------------
    def distribute(self, positions_raw, values, reset_hashmap=True):
        n, d = positions_raw.shape
        v = values.shape[1]
        self.m_positions = positions_raw
        if self.table is None:
            self.table = RefTable(self.cap, d, v)
        if self.m_values is None:
            self.m_values = self.table.values
        distributed = torch.zeros((n * (d + 1), d + v + 1), dtype=torch.float32, device="cuda")
        idx = torch.empty((n * (d + 1),), dtype=torch.int32, device="cuda").fill_(-1)
        w = torch.empty((n * (d + 1),), dtype=torch.float32, device="cuda").fill_(-1)
        new = self.clone_lattice()
        new.table = RefTable.__new__(RefTable)
        new.table.capacity, new.table.pos_dim = self.cap, d
        new.table.keys = self.table.keys.clone()          # the three clones of Lattice.cu:377-381
        new.table.entries = self.table.entries.clone()
        new.table.values = self.table.values.clone()
        new.table.nr_filled = self.table.nr_filled.clone()
        new.table.clear()
        new.m_values = new.table.values
        positions = positions_raw / self.sigmas_t
        self.k.launch(f"distribute<{d},{v}>", n, [_p(positions), _p(values), ctypes.c_int(n), _p(idx), _p(w), _p(distributed), new.table.struct()])
        new.dirty = True
        return new, distributed, idx, w

    # -- coarse vertices (Lattice.cu:706-740) -----------------------------------------------------------
    def create_coarse_verts_naive(self, positions_raw):
        d = self.pos_dim()
        c = RefHandle(self.cap, [s * 2.0 for s in self.m_sigmas], self.m_lvl + 1)
        c.m_positions = self.m_positions
        c.table = RefTable(self.cap, d)                    # zeros allocs + clear
        c.table.clear()                                    # begin_splat() clears a second time
        c.m_values = torch.zeros((1, self.val_dim()), dtype=torch.float32, device="cuda")
        positions = positions_raw / c.sigmas_t
        n = positions_raw.shape[0]
        self.k.launch(f"kernel_splat<{d},1>", n, [_p(positions), ctypes.c_int(n), _p(None), _p(None), c.table.struct(), ctypes.c_bool(False)])
        return c

    # -- im2row / row2im / conv (Lattice.cu:424-474, 612-667) ----------------------------------------------
    def im2row(self, nbrs, filter_extent, dilation, flip):
        nbrs = self if nbrs is None else nbrs
        nv = self.nr_lattice_vertices()
        d, v = self.pos_dim(), nbrs.val_dim()
        rowified = torch.zeros((nv, filter_extent * v), dtype=torch.float32, device="cuda")
        self.k.launch(f"im2row<{d},{v}>", nv,
                      [ctypes.c_int(nv), _p(rowified), ctypes.c_int(filter_extent), ctypes.c_int(dilation), self._struct(),
                       nbrs._struct(), ctypes.c_int(self.m_lvl), ctypes.c_int(nbrs.m_lvl), ctypes.c_bool(flip), ctypes.c_bool(False)])
        return rowified

    def row2im(self, rowified, dilation, filter_extent, nr_filters, nbrs=None):
        nbrs = self if nbrs is None else nbrs
        nv = self.nr_lattice_vertices()
        d, v = self.pos_dim(), self.val_dim()
        self.m_values = torch.zeros((nv, v), dtype=torch.float32, device="cuda")
        self.k.launch(f"row2im<{d},{v}>", self.cap,
                      [ctypes.c_int(self.cap), _p(rowified), ctypes.c_int(filter_extent), ctypes.c_int(dilation), self._struct(),
                       nbrs._struct(), ctypes.c_int(self.m_lvl), ctypes.c_int(nbrs.m_lvl), ctypes.c_bool(False)])
        return self.m_values

    def convolve_im2row_standalone(self, filter_bank, dilation, nbrs=None, flip=False, bias=None):
        nbrs = self if nbrs is None else nbrs
        F = filter_bank.shape[0] // nbrs.val_dim()
        rowified = self.im2row(nbrs, F, dilation, flip)
        out = self.clone_lattice()                    # new handle => its count is "dirty" => D2H sync (Lattice.cu:470)
        out.m_values = rowified.mm(filter_bank.contiguous())
        out.nr_lattice_vertices()
        return out

    def conv_weight_grad(self, nbrs, grad_values, filter_extent, dilation):
        # lattice_funcs.py:298-302: re-materialise im2row, transpose, mm
        return self.im2row(nbrs, filter_extent, dilation, False).transpose(0, 1).mm(grad_values)

    @staticmethod
    def filter_for_data_grad(filter_bank, filter_extent, val_dim):
        # lattice_funcs.py:304-311
        nr_filters = filter_bank.shape[1]
        fb = filter_bank.transpose(0, 1).view(nr_filters, filter_extent, val_dim).transpose(0, 1).contiguous()
        return fb.reshape(filter_extent * nr_filters, val_dim)

    # -- slice family (Lattice.cu:878-1142) ----------------------------------------------------------------
    def gather_standalone_with_precomputation(self, positions_raw, idx, w):
        n, d = positions_raw.shape
        v = self.val_dim()
        out = torch.zeros((n, (d + 1) * (v + 1)), dtype=torch.float32, device="cuda")
        positions = positions_raw / self.sigmas_t
        self.k.launch(f"gather_with_precomputation<{d},{v}>", n, [_p(positions), _p(out), ctypes.c_int(n), _p(idx), _p(w), self._struct()])
        return out

    def gather_backwards_standalone_with_precomputation(self, positions_raw, grad, idx, w):
        n, d = positions_raw.shape
        v = grad.shape[1] // (d + 1) - 1
        self.m_values = torch.zeros((self.nr_lattice_vertices(), v), dtype=torch.float32, device="cuda")
        self.k.launch(f"gather_backwards_with_precomputation<{d},{v}>", n, [ctypes.c_int(n), _p(grad), _p(idx), _p(w), self._struct()])

    def slice_classify_with_precomputation(self, positions_raw, dw, cw, cb, nc, idx, w):
        n, d = positions_raw.shape
        v = self.val_dim()
        out = torch.zeros((n, nc), dtype=torch.float32, device="cuda")
        positions = positions_raw / self.sigmas_t
        self.k.launch(f"slice_classify_with_precomputation<{d},{v},{nc}>", n,
                      [_p(positions), _p(out), _p(dw.contiguous()), _p(cw.contiguous()), _p(cb.contiguous()), ctypes.c_int(n), _p(idx), _p(w), self._struct()])
        return out

    def slice_classify_backwards_with_precomputation(self, g, positions_raw, init_vals, dw, cw, cb, nc, g_lv, g_dw, g_w, g_b, idx, w):
        n, d = positions_raw.shape
        v = init_vals.shape[1]
        self.k.launch(f"slice_classify_backwards_with_precomputation<{d},{v},{nc}>", n,
                      [ctypes.c_int(n), _p(g), _p(init_vals.contiguous()), _p(idx), _p(w), _p(dw.contiguous()), _p(cw.contiguous()),
                       _p(cb.contiguous()), _p(g_lv), _p(g_dw), _p(g_w), _p(g_b), self._struct(init_vals)])


# --------------------------------------------------------------------------------------------------
# import shims for what the reference's Python imports but this image does not have
def _scatter_max(src, index, dim=0, dim_size=None):
    """torch_scatter.scatter_max(src, index, dim=0): (max per index, row attaining it); library scatter ops only."""
    m, c = src.shape
    nv = int(index.max().item()) + 1 if dim_size is None else dim_size       # torch_scatter sizes the output the same way (one sync)
    idx = index.long().unsqueeze(1).expand(-1, c)
    out = torch.full((nv, c), float("-inf"), device=src.device).scatter_reduce(0, idx, src, "amax", include_self=True)
    hit = src == out.gather(0, idx)
    rows = torch.where(hit, torch.arange(m, device=src.device).unsqueeze(1).expand(-1, c), torch.full_like(idx, m))
    arg = torch.full((nv, c), m, dtype=torch.int64, device=src.device).scatter_reduce(0, idx, rows, "amin", include_self=True)
    out = src.gather(0, arg.clamp(max=m - 1))          # differentiable w.r.t. src like scatter_max
    return out, arg.clamp(max=m - 1)


def _scatter_add(src, index, dim=0, dim_size=None):
    nv = int(index.max().item()) + 1 if dim_size is None else dim_size
    shape = (nv,) + tuple(src.shape[1:])
    return torch.zeros(shape, device=src.device, dtype=src.dtype).index_add_(0, index.long(), src)


def _scatter_mean(src, index, dim=0, dim_size=None):
    s = _scatter_add(src, index, dim, dim_size)
    cnt = _scatter_add(torch.ones(index.shape[0], device=src.device), index, 0, s.shape[0]).clamp(min=1)
    return s / cnt.view(-1, *([1] * (s.dim() - 1)))


def install_shims():
    import sys
    import types
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref_py = os.path.join(here, "baseline", "_ref")
    if not os.path.isfile(os.path.join(ref_py, "latticenet_py", "lattice", "models.py")):
        raise RuntimeError("baseline/_ref/latticenet_py is missing: run `python -m oracle.install_ref_py` where /root/reference is mounted")
    if ref_py not in sys.path:
        sys.path.insert(0, ref_py)

    class _Profiler:
        @staticmethod
        def is_profiling_gpu():
            return False

        @staticmethod
        def start(name):
            pass

        @staticmethod
        def end(name):
            pass

    easypbr = types.ModuleType("easypbr")
    easypbr.Profiler = _Profiler
    easypbr.Mesh = type("Mesh", (), {})
    easypbr.Scene = type("Scene", (), {})
    easypbr.__all__ = ["Profiler", "Mesh", "Scene"]
    sys.modules.setdefault("easypbr", easypbr)
    latticenet = types.ModuleType("latticenet")
    latticenet.Lattice = RefHandle
    latticenet.HashTable = RefTable
    sys.modules.setdefault("latticenet", latticenet)
    ts = types.ModuleType("torch_scatter")
    ts.scatter_max, ts.scatter_add, ts.scatter_mean = _scatter_max, _scatter_add, _scatter_mean
    sys.modules.setdefault("torch_scatter", ts)
    tc = types.ModuleType("termcolor")
    tc.colored = lambda text, *a, **k: text
    sys.modules.setdefault("termcolor", tc)


def build_reference_model(nr_classes, model_params, pos_dim=3):
    """The reference's own LNN (latticenet_py/lattice/models.py), imported unmodified."""
    import contextlib
    import io
    install_shims()
    RefKernels.get()
    RefHandle.m_expected_pos_dim = pos_dim
    from latticenet_py.lattice.models import LNN as RefLNN      # the reference's file
    with contextlib.redirect_stdout(io.StringIO()):              # its constructor prints one line per block
        model = RefLNN(nr_classes, model_params)
    return model


def run(args, cloud_fn, cfg):
    from lattice_net_b200.params import ModelParams              # plain .cfg value holder with the reference's accessor names
    install_shims()
    from latticenet_py.lattice.lovasz_loss import LovaszSoftmax # the reference's loss

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    model = build_reference_model(cfg["nr_classes"], ModelParams())
    pool = 16
    clouds = [cloud_fn(i) for i in range(pool)]
    dev_clouds = [(torch.from_numpy(p).to(dev), torch.zeros((cfg["nr_points"], 1), device=dev), torch.from_numpy(l).to(dev)) for p, l in clouds]

    lattice = RefHandle(cfg["capacity"], [cfg["sigma"]] * 3)
    with torch.no_grad():
        model(lattice, *dev_clouds[0][:2])                       # lazily created layers (the reference creates its optimizer after this too)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=3e-4, amsgrad=True)       # ln_train.py:163-165
    loss_fn = LovaszSoftmax(ignore_index=-100)                   # ln_train.py:128-130, 156-158
    secondary_fn = torch.nn.NLLLoss(ignore_index=-100)
    flush = torch.empty((256 << 20) // 4, dtype=torch.float32, device=dev)

    def step(i):
        pos, vals, labels = dev_clouds[i % pool]
        logsm, _ = model(lattice, pos, vals)
        loss = 0.5 * loss_fn(logsm, labels) + 0.5 * secondary_fn(logsm, labels)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    # GPU-busy share: kernel time of one step (a profiler pass outside the timed region) against the step's wall time
    busy_ms = None
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step(0)
            torch.cuda.synchronize()
        busy_ms = sum(e.device_time_total for e in prof.key_averages()) * 1e-3
    except Exception:
        pass
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        flush.fill_(float(i))
        loss = step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    value = args.steps / (ms * 1e-3)
    scenes = []
    if not getattr(args, "no_extras", False):
        # the scene-sized scans of BASELINE configs[2] / [3] through the same arm (eager, as the reference runs)
        import bench_scenes
        del model, opt
        torch.cuda.empty_cache()
        for name in ("kitti", "scannet"):
            try:
                r = bench_scenes.run_scene(name, "reference", 3, 1, 1, "eager")
                scenes.append({k: r[k] for k in ("scene", "n_points", "vertices_per_level", "fwd_bwd_ms", "scans_per_s", "points_per_s", "inference_ms",
                                                  "inference_points_per_s", "execution", "conv")})
            except Exception as exc:
                scenes.append({"scene": name, "error": f"{type(exc).__name__}: {exc}"})
            torch.cuda.empty_cache()
    return {
        "scenes": scenes,
        "impl": "reference", "metric": "scans/sec fwd+bwd", "value": value, "unit": "scans/s", "n_gpus": int(getattr(args, "gpus", 1) or 1), "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "LatticeNet ShapeNet-part segmentation fwd+bwd+AdamW (lnn_train_shapenet.cfg arch), 2048-pt synthetic clouds, 1 scene/step/GPU",
                   "nr_points": cfg["nr_points"], "nr_classes": cfg["nr_classes"], "sigma": cfg["sigma"], "hash_table_capacity": cfg["capacity"],
                   "arm": "reference Python modules unmodified (baseline/_ref/latticenet_py: lattice_funcs, lattice_modules, models, LovaszSoftmax, utils) + the "
                          "reference's own CUDA kernels (NVRTC build of the unmodified LatticeGPU.cuh, driver-JITed on this GPU) under a restatement of its C++ "
                          "host class (im2row buffer + fp32 mm, C-sized clears, per-handle D2H syncs); eager, torch.optim.AdamW(amsgrad) unfused as upstream; "
                          "the reference has no CPU implementation",
                   "gpu_busy_ms_per_step": busy_ms, "gpu_busy_share_of_step": (busy_ms / (ms / args.steps)) if busy_ms else None,
                   "l2": "flushed between steps by a 256 MiB write (inside the timed region)",
                   "ranks_used": "1 (the reference is single-process / single-GPU; under torchrun rank 0 alone runs it)"},
        "cpu_baseline": {"value": value, "unit": "scans/s", "cores": 0, "kind": "reference",
                         "sample": f"{args.steps} full training steps on the B200 (the reference is CUDA-only, README.md:20; no host-core implementation exists)"},
        "e2e": {"value": value, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "final_loss": float(loss.item()),
    }
