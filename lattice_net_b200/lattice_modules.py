"""nn.Module layer over the lattice Functions -- the class names, constructor arguments, forward
signatures and parameter names (`weight`, `bias`, `weight_g`, `weight_v`, ...) of
/root/reference/latticenet_py/lattice/lattice_modules.py, so checkpoints and model code written
against the reference keep working.

Differences in mechanism (not in results):
  * `ConvLatticeIm2RowModule` runs the fused implicit-GEMM convolution instead of
    `Im2RowLattice.apply(...)` followed by `mm` (lattice_modules.py:240-242);
  * the `torch_scatter` calls (lattice_modules.py:78,688,692) are replaced by the segmented-reduction
    kernels of the C ABI (`ln_scatter_max`, `ln_scatter_sum_count`);
  * modules create their parameters on the device of the lattice they first see, not on "cuda:0".
"""
import math
import sys

import numpy as np
import torch
from torch.nn import functional as F

from . import lattice as _lattice
from ._cabi import call, ptr, stream_ptr
from .lattice import Lattice
from .lattice_funcs import (CoarsenLattice, ConvIm2RowLattice, DistributeLattice, ExpandLattice, FinefyLattice,
                            GatherLattice, Im2RowLattice, SliceClassifyLattice, SliceLattice, SplatLattice)


def _default_device():
    return torch.device("cuda", torch.cuda.current_device())


# --------------------------------------------------------------------------------------------------
# segmented reductions over "points that share a vertex" (torch_scatter replacements)
class _ScatterMax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, index_i32, nv):
        src = src.contiguous()
        m, c = src.shape
        out = torch.empty((nv, c), dtype=torch.float32, device=src.device)
        arg = torch.empty((nv, c), dtype=torch.int32, device=src.device)
        work = torch.empty((nv, c), dtype=torch.int64, device=src.device)
        call("ln_scatter_max", ptr(src), ptr(index_i32), m, c, nv, ptr(out), ptr(arg), ptr(work), stream_ptr(src.device))
        ctx.save_for_backward(arg)
        ctx.m = m
        ctx.mark_non_differentiable(arg)
        return out, arg

    @staticmethod
    def backward(ctx, grad_out, _grad_arg):
        (arg,) = ctx.saved_tensors
        nv, c = arg.shape
        # one padding row (index m) swallows the empty vertices
        grad_src = grad_out.new_zeros((ctx.m + 1, c))
        grad_src.scatter_(0, arg.long(), grad_out)
        return grad_src[:ctx.m], None, None


def scatter_max(src, index, nv):
    """torch_scatter.scatter_max(src, index, dim=0) with dim_size nv: (max [nv,c], argmax [nv,c])."""
    return _ScatterMax.apply(src, index.to(torch.int32).contiguous(), int(nv))


def scatter_sum_count(src, index, nv):
    """(sum [nv,c], count [nv]) of the rows of src that map to each vertex; no autograd (used under no_grad)."""
    src = src.contiguous()
    m, c = src.shape
    out = torch.empty((nv, c), dtype=torch.float32, device=src.device)
    cnt = torch.empty((nv,), dtype=torch.float32, device=src.device)
    call("ln_scatter_sum_count", ptr(src), ptr(index.to(torch.int32).contiguous()), m, c, int(nv), ptr(out), ptr(cnt), stream_ptr(src.device))
    return out, cnt


_GN_SMALL_ROWS = 2048     # lattices up to this many rows run the one-CTA-per-group kernels (ln_norm.cu: kGnRows * kGnThreads)


def _gn_workspace(nv, c, groups, device):
    """Scratch for the row-tiled GroupNorm kernels of scene-sized lattices (None when the library needs none).  A
    fresh tensor per call: the caching allocator recycles it in stream order, and it stays valid inside a CUDA graph."""
    if nv <= _GN_SMALL_ROWS and (c // groups) in (1, 2, 3, 4, 6, 8):
        return None
    from ._cabi import load
    nbytes = int(load().ln_group_norm_workspace_bytes(int(nv), int(c), int(groups)))
    return torch.empty((nbytes // 4,), dtype=torch.float32, device=device) if nbytes > 0 else None


def _gn_forward(x, gamma, beta, groups, eps, relu, nv_dev):
    x = x.contiguous()
    nv, c = x.shape
    y = torch.empty_like(x)
    stats = torch.empty((groups, 2), dtype=torch.float32, device=x.device)
    call("ln_group_norm_fwd", ptr(x), ptr(gamma.contiguous()), ptr(beta.contiguous()), nv, ptr(nv_dev), c, groups, float(eps),
         1 if relu else 0, ptr(y), ptr(stats), ptr(_gn_workspace(nv, c, groups, x.device)), stream_ptr(x.device))
    return x, y, stats


def _gn_backward(dy, x, y, gamma, beta, stats, groups, relu, nv_dev, dx_add):
    """-> (dx [+ dx_add], dgamma, dbeta); the affine gradients land in their gradient-bucket slices when a bucket is active."""
    nv, c = x.shape
    dx = torch.empty_like(x)
    dgamma = _lattice.grad_target(gamma)
    dbeta = _lattice.grad_target(beta)
    if dgamma is None or dbeta is None:
        dgamma, dbeta = torch.empty_like(gamma), torch.empty_like(gamma)
    if dx_add is not None:
        dx_add = dx_add.contiguous()
    call("ln_group_norm_bwd", ptr(dy.contiguous()), ptr(x), ptr(y), ptr(gamma.contiguous()), ptr(stats), ptr(dx_add), nv, ptr(nv_dev), c,
         groups, 1 if relu else 0, ptr(dx), ptr(dgamma), ptr(dbeta), ptr(_gn_workspace(nv, c, groups, x.device)), stream_ptr(x.device))
    return dx, dgamma, dbeta


class _GroupNormReLU(torch.autograd.Function):
    """GroupNorm (+ReLU) on vertex-major values [nv x C] in one kernel each way (ln_group_norm_fwd/bwd)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, groups, eps, relu, nv_dev=None):
        # nv_dev: device int32[1] with the actual vertex count when x is padded to a row bound (static-shape mode)
        x, y, stats = _gn_forward(x, gamma, beta, groups, eps, relu, nv_dev)
        ctx.save_for_backward(x, y, gamma, beta, stats)
        ctx.groups, ctx.relu, ctx.nv_dev = groups, relu, nv_dev
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, gamma, beta, stats = ctx.saved_tensors
        dx, dgamma, dbeta = _gn_backward(dy, x, y, gamma, beta, stats, ctx.groups, ctx.relu, ctx.nv_dev, None)
        return dx, dgamma, dbeta, None, None, None, None


class _GroupNormReLUSplit(torch.autograd.Function):
    """The head of a residual block: returns (GroupNorm(+ReLU)(x), x).  Both users of x -- the normalised branch and
    the skip connection -- hang off this one node, so the two gradients meet INSIDE the backward kernel
    (dx = gn_backward(dy) + d_skip, ln_group_norm_bwd's dx_add) instead of in an add kernel launched by autograd
    (lattice_modules.py:1255-1290 / 1322-1358: `identity = lv ... lv = lv + identity`)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, groups, eps, relu, nv_dev=None):
        xc, y, stats = _gn_forward(x, gamma, beta, groups, eps, relu, nv_dev)
        ctx.save_for_backward(xc, y, gamma, beta, stats)
        ctx.groups, ctx.relu, ctx.nv_dev = groups, relu, nv_dev
        return y, x.view_as(x)

    @staticmethod
    def backward(ctx, dy, d_skip):
        x, y, gamma, beta, stats = ctx.saved_tensors
        if dy is None:
            dy = torch.zeros_like(x)
        dx, dgamma, dbeta = _gn_backward(dy, x, y, gamma, beta, stats, ctx.groups, ctx.relu, ctx.nv_dev, d_skip)
        return dx, dgamma, dbeta, None, None, None, None


class _LinearFn(torch.autograd.Function):
    """y = x W^T (+ bias) (+ residual) for the 1x1 layers (lattice_modules.py:806-832 uses torch.nn.Linear): forward, data
    gradient and weight gradient run through the lattice-convolution kernels with filter extent 1 -- tcgen05 3xTF32 when
    in_features % 32 == 0 -- with the bias / skip connection folded into the epilogue."""

    @staticmethod
    def forward(ctx, x, weight, bias=None, residual=None):
        x = x.contiguous()
        ctx.save_for_backward(x, weight)
        ctx.has_bias, ctx.has_residual = bias is not None, residual is not None
        return _lattice.linear_forward(x, weight, bias, residual)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        grad_param = weight if (weight.is_leaf and weight.requires_grad) else None
        dx, dw = _lattice.linear_backward(x, weight, dy, ctx.needs_input_grad[0], grad_param)
        db = dy.sum(0) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return dx, dw, db, (dy if (ctx.has_residual and ctx.needs_input_grad[3]) else None)


class _PointNetFusedFn(torch.autograd.Function):
    """(positions, values, indices, weights, sigmas, nv_rows, quirk, v0, g0, b0, v1, g1, b1, v2, g2, b2) -> [nv_rows x 2*64]"""

    @staticmethod
    def forward(ctx, positions, values, indices, weights, sigmas, nv_rows, quirk, *params):
        import ctypes
        from ._cabi import load
        lib = load()
        n, d = positions.shape
        v = values.shape[1]
        dev = positions.device
        widths = [int(params[3 * l].shape[0]) for l in range(3)]
        scratch = _lattice._zeroed(1, int(lib.ln_pointnet_scratch_floats(d, int(nv_rows), widths[2])), dev).view(-1)
        out = torch.empty((nv_rows, 2 * widths[2]), dtype=torch.float32, device=dev)
        arg = torch.empty((nv_rows, widths[2]), dtype=torch.int32, device=dev)
        ps = [p.contiguous() for p in params]
        ptrs = (ctypes.c_void_p * 9)(*[p.data_ptr() for p in ps])
        call("ln_pointnet_fwd", ptr(positions), ptr(sigmas), ptr(values), ptr(indices), ptr(weights), n, d, v, ptrs, widths[0], widths[1],
             widths[2], int(nv_rows), 1 if quirk else 0, 4, ptr(scratch), ptr(out), ptr(arg), stream_ptr(dev))
        ctx.save_for_backward(positions, values, indices, sigmas, scratch, arg, *ps)
        ctx.quirk, ctx.widths = bool(quirk), widths
        ctx.mark_non_differentiable(arg)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        import ctypes
        from ._cabi import load
        positions, values, indices, sigmas, scratch, arg = ctx.saved_tensors[:6]
        ps = ctx.saved_tensors[6:]
        n, d = positions.shape
        v = values.shape[1]
        dev = positions.device
        w = ctx.widths
        gscratch = _lattice._zeroed(1, int(load().ln_pointnet_grad_scratch_floats(d, v, w[0], w[1], w[2])), dev).view(-1)
        grads = []
        for p in ps:
            t = _lattice.grad_target(p)
            grads.append(t if t is not None else torch.empty_like(p))
        ptrs = (ctypes.c_void_p * 9)(*[p.data_ptr() for p in ps])
        gptrs = (ctypes.c_void_p * 9)(*[g.data_ptr() for g in grads])
        call("ln_pointnet_bwd", ptr(positions), ptr(sigmas), ptr(values), ptr(indices), n, d, v, ptrs, gptrs, w[0], w[1], w[2],
             1 if ctx.quirk else 0, ptr(scratch), ptr(grad_out.contiguous()), ptr(arg), ptr(gscratch), stream_ptr(dev))
        return (None,) * 7 + tuple(grads)


class _DeltaWFn(torch.autograd.Function):
    """Learned barycentric offsets of the DeformSlice head in one kernel each way (ln_deltaw_fwd / _bwd):
    (bottleneck values [nv x 8], indices, weights, gamma [9], beta [9], lin_w [1 x 9], lin_b [1], pos_dim) -> [N x (d+1)]."""

    @staticmethod
    def forward(ctx, values, indices, weights, gamma, beta, lin_w, lin_b, pos_dim):
        values = values.contiguous()
        n = indices.numel() // (pos_dim + 1)
        out = torch.empty((n, pos_dim + 1), dtype=torch.float32, device=values.device)
        call("ln_deltaw_fwd", ptr(values), ptr(indices), ptr(weights), ptr(gamma.contiguous()), ptr(beta.contiguous()), ptr(lin_w.contiguous()),
             ptr(lin_b.contiguous()), n, pos_dim, int(values.shape[1]), ptr(out), stream_ptr(values.device))
        ctx.save_for_backward(values, indices, weights, gamma, beta, lin_w, lin_b)
        ctx.pos_dim = pos_dim
        return out

    @staticmethod
    def backward(ctx, grad_out):
        values, indices, weights, gamma, beta, lin_w, lin_b = ctx.saved_tensors
        n = indices.numel() // (ctx.pos_dim + 1)
        dev = values.device
        g_values = _lattice._zeroed(values.shape[0], values.shape[1], dev)
        grads = []
        for p in (lin_w, lin_b, gamma, beta):        # bucket slices when a zeroed gradient bucket is active, fresh zeros otherwise
            t = _lattice.grad_target(p)
            grads.append(t)
        if any(t is None for t in grads):
            small = torch.zeros((28,), dtype=torch.float32, device=dev)
            grads = [small[0:9].view_as(lin_w), small[9:10].view_as(lin_b), small[10:19].view_as(gamma), small[19:28].view_as(beta)]
        call("ln_deltaw_bwd", ptr(values), ptr(indices), ptr(weights), ptr(gamma.contiguous()), ptr(beta.contiguous()), ptr(lin_w.contiguous()),
             ptr(grad_out.contiguous()), n, ctx.pos_dim, int(values.shape[1]), ptr(g_values), ptr(grads[0]), ptr(grads[1]), ptr(grads[2]),
             ptr(grads[3]), stream_ptr(dev))
        return g_values, None, None, grads[2], grads[3], grads[0], grads[1], None


def linear(x, weight, bias=None, residual=None):
    """F.linear(x, weight, bias) (+ residual) on the lattice kernels (CUDA fp32 2-D inputs), torch elsewhere."""
    if x.is_cuda and x.dim() == 2 and x.dtype == torch.float32 and weight.dtype == torch.float32:
        return _LinearFn.apply(x, weight, bias, residual)
    y = F.linear(x, weight, bias)
    return y if residual is None else y + residual


# None: GroupNorm(+ReLU) always runs the kernels of ln_norm.cu (one CTA per group for ShapeNet-sized lattices, row-tiled
# over all SMs for scene-sized ones).  An integer restores torch.nn.GroupNorm above that many elements per group
# (0 = always torch: what bench.py's reference arm sets).
FUSED_NORM_MAX_ELEMS_PER_GROUP = None

# The reference treats lattice vertex 0 as the "invalid" row: points whose index was clamped from -1 land there,
# so its mean / features are zeroed (lattice_modules.py:72-94, 683, 712).  Which real vertex gets id 0 is a race
# of the hash insert, in the reference as here, which makes model outputs differ from run to run.  True keeps the
# reference behaviour; tests that compare two independent runs switch it off to get a numbering-invariant model.
REFERENCE_VERTEX0_QUIRK = True


# --------------------------------------------------------------------------------------------------
# weight normalisation with a per-output gain and a whole-tensor norm: what the reference's
# weight_norm_wrapper(cls, g_dim, v_dim=None) computes (latticenet_py/utils/utils.py:72-158)
def _split_weight_norm(module, g_dim):
    w = module.weight
    del module._parameters["weight"]
    shape = [1] * w.dim()
    shape[g_dim] = w.shape[g_dim]
    module.weight_g = torch.nn.Parameter(torch.full(shape, float(w.detach().norm()), device=w.device))
    module.weight_v = torch.nn.Parameter(w.detach().clone())
    module._wn_g_dim = g_dim


class _WeightNormFn(torch.autograd.Function):
    """w = v * (g / ||v||) in one kernel each way (ln_weight_norm_fwd / _bwd) for 2-D CUDA weights."""

    @staticmethod
    def forward(ctx, v, g, g_dim):
        v, g = v.contiguous(), g.contiguous()
        w = torch.empty_like(v)
        call("ln_weight_norm_fwd", ptr(v), ptr(g), int(v.shape[0]), int(v.shape[1]), 1 if g_dim == 1 else 0, ptr(w), stream_ptr(v.device))
        ctx.save_for_backward(v, g)
        ctx.g_dim = g_dim
        return w

    @staticmethod
    def backward(ctx, dw):
        v, g = ctx.saved_tensors
        dv = _lattice.grad_target(v)
        dg = _lattice.grad_target(g)
        if dv is None or dg is None:
            dv, dg = torch.empty_like(v), torch.empty_like(g)
        call("ln_weight_norm_bwd", ptr(v), ptr(g), ptr(dw.contiguous()), int(v.shape[0]), int(v.shape[1]), 1 if ctx.g_dim == 1 else 0, ptr(dv),
             ptr(dg), stream_ptr(v.device))
        return dv, dg, None


def _normed_weight(module):
    v, g = module.weight_v, module.weight_g
    if v.is_cuda and v.dim() == 2 and v.dtype == torch.float32 and v.numel() <= (1 << 22):
        return _WeightNormFn.apply(v, g, module._wn_g_dim)
    return v * (g / v.norm())


class LinearWN(torch.nn.Linear):
    """Linear with weight norm (g_dim=0), reference: utils.LinearWN."""

    def __init__(self, in_features, out_features, bias=True):
        super().__init__(in_features, out_features, bias=bias)
        _split_weight_norm(self, 0)

    def forward(self, x):
        return F.linear(x, _normed_weight(self), self.bias)


def leaky_relu_init_(weight, fan_sum, alpha=0.2):
    gain = math.sqrt(2.0 / (1.0 + alpha ** 2))
    std = gain * math.sqrt(2.0 / fan_sum)
    with torch.no_grad():
        weight.uniform_(-std * math.sqrt(3.0), std * math.sqrt(3.0))


# --------------------------------------------------------------------------------------------------
class DropoutLattice(torch.nn.Module):
    def __init__(self, prob):
        super().__init__()
        self.dropout = torch.nn.Dropout2d(p=prob)

    def forward(self, lv):
        if lv.dim() != 2:
            sys.exit("the lattice values must be two dimensional: nr_lattice_vertices x val_dim")
        # channel-wise dropout: treat the channels as feature maps
        x = lv.t().unsqueeze(0).unsqueeze(3)
        return self.dropout(x).squeeze(3).squeeze(0).t()


class SplatLatticeModule(torch.nn.Module):
    def forward(self, lattice_py, positions, values):
        lv, ls_wrap, indices, weights = SplatLattice.apply(lattice_py, positions, values)
        return lv, ls_wrap.lattice, indices, weights


class DistributeLatticeModule(torch.nn.Module):
    """distribute + subtraction of the per-vertex mean position (lattice_modules.py:52-96)."""

    def forward(self, lattice, positions, values, reset_hashmap=True):
        wrap, distributed, indices, weights = DistributeLattice.apply(lattice, positions, values, reset_hashmap)
        dist_lattice = wrap.lattice
        d = positions.shape[1]
        idx = indices.clamp(min=0)
        nv = dist_lattice.nr_lattice_vertices()
        pos_cols = distributed[:, :d].contiguous()
        sums, counts = scatter_sum_count(pos_cols, idx, nv)
        mean = sums / counts.clamp(min=1.0).unsqueeze(1)
        idx_long = idx.long()
        if REFERENCE_VERTEX0_QUIRK:
            mean[0] = 0.0                              # vertex 0 doubles as the "invalid" row in the reference
        distributed[:, :d] = pos_cols - mean.index_select(0, idx_long)
        if REFERENCE_VERTEX0_QUIRK:
            distributed = distributed.masked_fill((idx_long == 0).unsqueeze(1), 0.0)
        return dist_lattice, distributed, indices, weights


class ExpandLatticeModule(torch.nn.Module):
    def __init__(self, point_multiplier, noise_stddev, expand_values):
        super().__init__()
        self.point_multiplier, self.noise_stddev, self.expand_values = point_multiplier, noise_stddev, expand_values

    def forward(self, lattice_values, lattice_structure, positions):
        lattice_structure.set_values(lattice_values)
        lv, ls_wrap = ExpandLattice.apply(lattice_values, lattice_structure, positions, self.point_multiplier,
                                          self.noise_stddev, self.expand_values)
        ls = ls_wrap.lattice
        ls.set_values(lv)
        return lv, ls


def _init_lattice_filter(weight, bias, fan_div=1.0, std_mul=1.0):
    # uniform with std = gain/sqrt(fan_out) (lattice_modules.py:199-213, 274-292)
    fan = torch.nn.init._calculate_correct_fan(weight, "fan_out") / fan_div
    std = torch.nn.init.calculate_gain("relu", 1) / math.sqrt(fan) * std_mul
    bound = math.sqrt(3.0) * std
    with torch.no_grad():
        weight.uniform_(-bound, bound)
        if bias is not None:
            _, fan_out = torch.nn.init._calculate_fan_in_and_fan_out(weight)
            b = 1.0 / math.sqrt(fan_out)
            bias.uniform_(-b, b)


class ConvLatticeModule(torch.nn.Module):
    """Lattice convolution whose input width is discovered at the first call (lattice_modules.py:120-171)."""

    def __init__(self, nr_filters, neighbourhood_size, dilation=1, bias=True):
        super().__init__()
        self.first_time = True
        self.weight = None
        self.bias = None
        self.neighbourhood_size, self.nr_filters, self.dilation, self.use_bias = neighbourhood_size, nr_filters, dilation, bias

    def forward(self, lattice_values, lattice_structure):
        lattice_structure.set_values(lattice_values)
        if self.first_time:
            self.first_time = False
            extent = lattice_structure.get_filter_extent(self.neighbourhood_size)
            dev = lattice_values.device
            self.weight = torch.nn.Parameter(torch.empty(extent * lattice_structure.val_dim(), self.nr_filters, device=dev))
            if self.use_bias:
                self.bias = torch.nn.Parameter(torch.empty(self.nr_filters, device=dev))
            _init_lattice_filter(self.weight, self.bias)
        lv, ls_wrap = ConvIm2RowLattice.apply(lattice_values, lattice_structure, self.weight, self.dilation)
        ls = ls_wrap.lattice
        if self.use_bias:
            lv = lv + self.bias
        ls.set_values(lv)
        return lv, ls


class _LatticeFilterModule(torch.nn.Module):
    """Common parameter handling of the conv / coarsen / finefy modules: `weight` is
    [filter_extent*in_channels x out_channels], row = slot*in_channels + channel."""
    _fan_div, _std_mul = 1.0, 1.0

    def __init__(self, in_channels, out_channels, bias, device=None):
        super().__init__()
        self.first_time = True
        self.in_channels, self.out_channels, self.use_bias = in_channels, out_channels, bias
        self.neighbourhood_size = 1
        self.filter_extent = Lattice.get_expected_filter_extent(self.neighbourhood_size)
        dev = device if device is not None else _default_device()
        self.weight = torch.nn.Parameter(torch.empty(self.filter_extent * in_channels, out_channels, device=dev))
        self.bias = torch.nn.Parameter(torch.empty(out_channels, device=dev)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        _init_lattice_filter(self.weight, self.bias, self._fan_div, self._std_mul)

    def _filter(self):
        return self.weight

    def _check_in(self, lattice):
        assert self.in_channels == lattice.val_dim(), \
            f"In channels doesn't match the val_dim of the lattice. In channels is {self.in_channels}, while val dim is {lattice.val_dim()}"


class ConvLatticeIm2RowModule(_LatticeFilterModule):
    def __init__(self, in_channels, out_channels, neighbourhood_size, dilation=1, bias=True, device=None):
        super().__init__(in_channels, out_channels, bias, device)
        self.neighbourhood_size, self.dilation = neighbourhood_size, dilation

    def forward(self, lattice_values, lattice_structure, residual=None):
        """residual (extension): [nv x out_channels] added to the result inside the convolution kernel."""
        lattice_structure.set_values(lattice_values)
        self._check_in(lattice_structure)
        if lattice_values.is_cuda and hasattr(lattice_structure, "conv_backward"):
            # bias and skip connection ride in the kernel's epilogue
            lv, ls_wrap = ConvIm2RowLattice.apply(lattice_values, lattice_structure, self._filter(), self.dilation,
                                                  self.bias if self.use_bias else None, residual)
            ls = ls_wrap.lattice   # a new handle: the value width may have changed
        else:
            lv, ls_wrap = ConvIm2RowLattice.apply(lattice_values, lattice_structure, self._filter(), self.dilation)
            ls = ls_wrap.lattice
            if self.use_bias:
                lv = lv + self.bias
            if residual is not None:
                lv = lv + residual
        ls.set_values(lv)
        return lv, ls


class CoarsenLatticeModule(_LatticeFilterModule):
    _fan_div, _std_mul = 2.0, 2.0   # lattice_modules.py:274-292

    def __init__(self, in_channels, out_channels, bias=False, device=None):
        super().__init__(in_channels, out_channels, bias, device)

    def forward(self, lattice_fine_values, lattice_fine_structure, coarsened_lattice=None):
        lattice_fine_structure.set_values(lattice_fine_values)
        self._check_in(lattice_fine_structure)
        lv, ls_wrap = CoarsenLattice.apply(lattice_fine_values, lattice_fine_structure, self._filter(), coarsened_lattice)
        ls = ls_wrap.lattice
        if self.use_bias:
            lv = lv + self.bias
        ls.set_values(lv)
        return lv, ls


class FinefyLatticeModule(_LatticeFilterModule):
    _fan_div, _std_mul = 2.0, 2.0   # lattice_modules.py:342-361

    def __init__(self, in_channels, out_channels, bias=False, device=None):
        super().__init__(in_channels, out_channels, bias, device)

    def forward(self, lattice_coarse_values, lattice_coarse_structure, lattice_fine_structure):
        lattice_coarse_structure.set_values(lattice_coarse_values)
        self._check_in(lattice_coarse_structure)
        lv, ls_wrap = FinefyLattice.apply(lattice_coarse_values, lattice_coarse_structure, lattice_fine_structure, self._filter())
        ls = ls_wrap.lattice
        if self.use_bias:
            lv = lv + self.bias
        ls.set_values(lv)
        return lv, ls


def _weight_normed(cls):
    """<cls> with weight norm on `weight` (g_dim=1): parameters `weight_g` [1 x out] and `weight_v`."""

    class Wrapped(cls):
        def __init__(self, *args, **kwargs):
            super().__init__(*args, **kwargs)
            _split_weight_norm(self, 1)

        def _filter(self):
            return _normed_weight(self)

    Wrapped.__name__ = cls.__name__.replace("Module", "WNModule")
    Wrapped.__qualname__ = Wrapped.__name__
    return Wrapped


ConvLatticeIm2RowWNModule = _weight_normed(ConvLatticeIm2RowModule)
CoarsenLatticeWNModule = _weight_normed(CoarsenLatticeModule)
FinefyLatticeWNModule = _weight_normed(FinefyLatticeModule)


class SliceLatticeModule(torch.nn.Module):
    def forward(self, lattice_values, lattice_structure, positions, splatting_indices=None, splatting_weights=None):
        lattice_structure.set_values(lattice_values)
        return SliceLattice.apply(lattice_values, lattice_structure, positions, splatting_indices, splatting_weights)


class GatherLatticeModule(torch.nn.Module):
    def forward(self, lattice_values, lattice_structure, positions, splatting_indices=None, splatting_weights=None):
        lattice_structure.set_values(lattice_values)
        if splatting_indices is None:   # the reference's 3-argument call cannot work (lattice_modules.py:406-409)
            gathered, _, _ = lattice_structure.gather_standalone_no_precomputation(positions)
            return gathered
        return GatherLattice.apply(lattice_values, lattice_structure, positions, splatting_indices, splatting_weights)


def filter_readings(model, with_backward=True):
    """Every (tensor, filter_extent, c_in, c_out, transposed) reading of the filter banks / 1x1 weights of `model` that a
    forward (and backward) pass will ask the tensor-core convolution for -- the input of lattice.prepare_filters()."""
    readings = []
    for m in model.modules():
        if isinstance(m, _LatticeFilterModule) and "weight" in m._parameters and m.weight is not None:
            w, fe, ci, co = m.weight, m.filter_extent, m.in_channels, m.out_channels
            readings.append((w, fe, ci, co, False))
            if with_backward:
                readings.append((w, fe, co, ci, True))
        elif isinstance(m, GnRelu1x1):
            w = m.linear.weight                       # [out x in]: the bank stored transposed
            co, ci = w.shape
            readings.append((w, 1, int(ci), int(co), True))
            if with_backward:
                readings.append((w, 1, int(co), int(ci), False))
    return readings


# --------------------------------------------------------------------------------------------------
class BatchNormLatticeModule(torch.nn.Module):
    def __init__(self, nr_params, affine=True, device=None):
        super().__init__()
        self.bn = torch.nn.BatchNorm1d(num_features=nr_params, momentum=0.1, affine=affine).to(device or _default_device())

    def forward(self, lattice_values, lattice_py):
        if lattice_values.dim() != 2:
            sys.exit("lattice should be 2 dimensional, nr_vertices x val_dim")
        lattice_values = self.bn(lattice_values)
        lattice_py.set_values(lattice_values)
        return lattice_values, lattice_py


class GroupNormLatticeModule(torch.nn.Module):
    def __init__(self, nr_params, affine=True, device=None):
        super().__init__()
        nr_groups = 32 if nr_params % 32 == 0 else int(nr_params / 2)   # lattice_modules.py:587-590
        self.gn = torch.nn.GroupNorm(nr_groups, nr_params).to(device or _default_device())

    def forward(self, lattice_values, lattice_py, do_set_values=True, relu=False, split=False):
        """split (extension): also return the input as the skip connection of a residual block -> (lv, skip, lattice);
        the two gradients are then summed inside the GroupNorm backward kernel."""
        if lattice_values.dim() != 2:
            sys.exit("lattice should be 2 dimensional, nr_vertices x val_dim")
        # statistics run over (channels of a group) x (all vertices): vertices are the "length" axis
        gn = self.gn
        nv, c = lattice_values.shape
        ht = getattr(lattice_py, "m_hash_table", None)      # (bench.py's reference arm passes its own handle type)
        st = ht.structure if ht is not None else None
        nv_dev = st.nr_filled if (st is not None and st.bound is not None) else None
        skip = lattice_values
        if lattice_values.is_cuda and (nv_dev is not None or FUSED_NORM_MAX_ELEMS_PER_GROUP is None
                                       or nv * (c // gn.num_groups) <= FUSED_NORM_MAX_ELEMS_PER_GROUP):
            if split:
                lv, skip = _GroupNormReLUSplit.apply(lattice_values, gn.weight, gn.bias, gn.num_groups, gn.eps, relu, nv_dev)
            else:
                lv = _GroupNormReLU.apply(lattice_values, gn.weight, gn.bias, gn.num_groups, gn.eps, relu, nv_dev)
        else:
            lv = gn(lattice_values.t().unsqueeze(0)).squeeze(0).t()
            if relu:
                lv = torch.relu(lv)
        if do_set_values:
            lattice_py.set_values(lv)
        if split:
            return lv, skip, lattice_py
        return lv, lattice_py

    def forward_relu(self, lattice_values, lattice_py, split=False):
        """GroupNorm followed by ReLU as one fused op (the GN -> ReLU -> conv pattern of every block)."""
        return self.forward(lattice_values, lattice_py, True, relu=True, split=split)


class PointNetModule(torch.nn.Module):
    """Per-point MLP, max-pool onto the vertices, one lattice conv (lattice_modules.py:620-733)."""

    def __init__(self, nr_output_channels_per_layer, nr_outputs_last_layer, device=None):
        super().__init__()
        self.first_time = True
        self.nr_output_channels_per_layer = list(nr_output_channels_per_layer)
        self.nr_outputs_last_layer = nr_outputs_last_layer
        self.layers = torch.nn.ModuleList([])
        self.act = torch.nn.LeakyReLU(0.2)
        self.last_conv = ConvLatticeIm2RowWNModule(in_channels=self.nr_output_channels_per_layer[-1] * 2,
                                                   out_channels=nr_outputs_last_layer, neighbourhood_size=1, dilation=1,
                                                   bias=True, device=device)
        fe = self.last_conv.filter_extent
        leaky_relu_init_(self.last_conv.weight_v, (self.last_conv.in_channels + nr_outputs_last_layer) * fe)
        with torch.no_grad():
            self.last_conv.weight_g.fill_(float(self.last_conv.weight_v.norm()))
            self.last_conv.bias.zero_()

    def init(self, distributed):
        if self.first_time:
            self.first_time = False
            nr_in = distributed.shape[1] - 1
            for nr_out in self.nr_output_channels_per_layer:
                lin = LinearWN(nr_in, nr_out, bias=True).to(distributed.device)
                leaky_relu_init_(lin.weight_v, nr_in + nr_out)
                with torch.no_grad():
                    lin.weight_g.fill_(float(lin.weight_v.norm()))
                    lin.bias.zero_()
                self.layers.append(lin)
                nr_in = nr_out

    def _init_layers(self, nr_in, device):
        if self.first_time:
            self.first_time = False
            for nr_out in self.nr_output_channels_per_layer:
                lin = LinearWN(nr_in, nr_out, bias=True).to(device)
                leaky_relu_init_(lin.weight_v, nr_in + nr_out)
                with torch.no_grad():
                    lin.weight_g.fill_(float(lin.weight_v.norm()))
                    lin.bias.zero_()
                self.layers.append(lin)
                nr_in = nr_out

    def fused_supported(self, pos_dim, val_dim):
        from ._cabi import load
        w = self.nr_output_channels_per_layer
        return len(w) == 3 and bool(load().ln_pointnet_supported(int(pos_dim), int(val_dim), int(w[0]), int(w[1]), int(w[2])))

    def forward_fused(self, lattice_py, positions, values, indices, weights):
        """distribute rows + mean subtraction + MLP + per-vertex max pooling + masks in three kernels, without the
        [N(d+1) x ...] tensors (csrc/ln_pointnet.cu); then the lattice convolution and activation as in forward()."""
        self._init_layers(positions.shape[1] + values.shape[1], positions.device)
        st = lattice_py.m_hash_table.structure
        nv_rows = lattice_py.nr_lattice_vertices()
        params = []
        for layer in self.layers:
            params += [layer.weight_v, layer.weight_g, layer.bias]
        reduced = _PointNetFusedFn.apply(positions.contiguous(), values.contiguous(), indices, weights, lattice_py._sigmas_on(st.device),
                                         nv_rows, REFERENCE_VERTEX0_QUIRK, *params)
        lattice_py.set_values(reduced)
        lv, ls = self.last_conv(reduced, lattice_py)
        lv = self.act(lv)
        ls.set_values(lv)
        return lv, ls

    def forward(self, lattice_py, distributed, indices):
        if self.first_time:
            self.init(distributed)
        barycentric = distributed[:, -1]
        x = distributed[:, :-1]
        for layer in self.layers:
            x = self.act(layer(x))
        idx = indices.clamp(min=0)
        nv = lattice_py.nr_lattice_vertices()
        reduced, argmax = scatter_max(x, idx, nv)
        _, counts = scatter_sum_count(torch.ones((idx.shape[0], 1), device=x.device), idx, nv)
        # barycentric weight of the point that won the max, per (vertex, channel)
        bary_pad = torch.cat([barycentric, barycentric.new_zeros(1)])
        bary_reduced = bary_pad.index_select(0, argmax.flatten().long()).view(argmax.shape)
        reduced = torch.cat((reduced, bary_reduced), 1)
        reduced = reduced.masked_fill((counts < 4).unsqueeze(1), 0.0)   # vertices touched by < 4 points
        if REFERENCE_VERTEX0_QUIRK:
            keep = torch.ones((nv, 1), device=x.device)
            keep[0] = 0.0                                                # row 0 collects the invalid points
            reduced = reduced * keep
        lattice_py.set_values(reduced)
        lv, ls = self.last_conv(reduced, lattice_py)
        lv = self.act(lv)
        ls.set_values(lv)
        return lv, ls


# --------------------------------------------------------------------------------------------------
class Conv1x1WN(torch.nn.Module):
    def __init__(self, in_channels, out_channels, bias, device=None):
        super().__init__()
        self.linear = LinearWN(in_channels, out_channels, bias=bias).to(device or _default_device())

    def forward(self, lv, ls):
        ls.set_values(lv)
        lv = self.linear(lv)
        ls.set_values(lv)
        return lv, ls


class Conv1x1WNAct(Conv1x1WN):
    def __init__(self, in_channels, out_channels, bias, device=None):
        super().__init__(in_channels, out_channels, bias, device)
        self.act = torch.nn.LeakyReLU(0.2)

    def forward(self, lv, ls):
        ls.set_values(lv)
        lv = self.act(self.linear(lv))
        ls.set_values(lv)
        return lv, ls


class Conv1x1(torch.nn.Module):
    def __init__(self, out_channels, bias):
        super().__init__()
        self.out_channels, self.use_bias, self.linear = out_channels, bias, None

    def forward(self, lv):
        if self.linear is None:
            self.linear = torch.nn.Linear(lv.shape[1], self.out_channels, bias=self.use_bias).to(lv.device)
            with torch.no_grad():
                torch.nn.init.kaiming_normal_(self.linear.weight, mode="fan_in", nonlinearity="relu")
        return self.linear(lv)


class GnRelu1x1(torch.nn.Module):
    def __init__(self, in_channels, out_channels, bias, device=None):
        super().__init__()
        dev = device or _default_device()
        self.norm = GroupNormLatticeModule(in_channels, device=dev)
        self.relu = torch.nn.ReLU(inplace=False)
        self.linear = torch.nn.Linear(in_channels, out_channels, bias=bias).to(dev)
        torch.nn.init.kaiming_normal_(self.linear.weight, mode="fan_in", nonlinearity="relu")

    def forward(self, lv, ls, residual=None, split=False):
        """residual / split (extensions): the skip connection of a bottleneck block enters / leaves here, see
        ConvLatticeIm2RowModule.forward and GroupNormLatticeModule.forward."""
        ls.set_values(lv)
        skip = None
        if split:
            lv, skip, ls = self.norm.forward_relu(lv, ls, split=True)
        else:
            lv, ls = self.norm.forward_relu(lv, ls)
        lv = linear(lv, self.linear.weight, self.linear.bias, residual)
        ls.set_values(lv)
        return (lv, skip, ls) if split else (lv, ls)


class Gn(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.norm = None

    def forward(self, lv, ls):
        ls.set_values(lv)
        if self.norm is None:
            self.norm = GroupNormLatticeModule(lv.shape[1], device=lv.device)
        lv, ls = self.norm(lv, ls)
        return lv, ls


class ConvAct(torch.nn.Module):
    def __init__(self, in_channels, out_channels, dilation, bias, with_dropout, device=None):
        super().__init__()
        self.conv = ConvLatticeIm2RowModule(in_channels=in_channels, out_channels=out_channels, neighbourhood_size=1,
                                            dilation=dilation, bias=bias, device=device)
        self.act = torch.nn.LeakyReLU(0.2)
        self.drop = DropoutLattice(0.2) if with_dropout else None

    def forward(self, lv, ls):
        ls.set_values(lv)
        if self.drop is not None:
            lv = self.drop(lv)
            ls.set_values(lv)
        lv_1, ls_1 = self.conv(lv, ls)
        lv_1 = self.act(lv_1)
        ls_1.set_values(lv_1)
        return lv_1, ls_1


class GnReluConv(torch.nn.Module):
    def __init__(self, in_channels, out_channels, dilation, bias, with_dropout, device=None):
        super().__init__()
        self.conv = ConvLatticeIm2RowModule(in_channels=in_channels, out_channels=out_channels, neighbourhood_size=1,
                                            dilation=dilation, bias=bias, device=device)
        self.norm = GroupNormLatticeModule(in_channels, device=device)
        self.relu = torch.nn.ReLU(inplace=False)
        self.drop = DropoutLattice(0.2) if with_dropout else None

    def forward(self, lv, ls, residual=None, split=False):
        ls.set_values(lv)
        skip = None
        if split:
            lv, skip, ls = self.norm.forward_relu(lv, ls, split=True)
        else:
            lv, ls = self.norm.forward_relu(lv, ls)
        if self.drop is not None:
            lv = self.drop(lv)
            ls.set_values(lv)
        lv_1, ls_1 = self.conv(lv, ls, residual) if residual is not None else self.conv(lv, ls)
        ls_1.set_values(lv_1)
        return (lv_1, skip, ls_1) if split else (lv_1, ls_1)


class BnReluConv(torch.nn.Module):
    def __init__(self, nr_filters, dilation, bias):
        super().__init__()
        self.conv = ConvLatticeModule(nr_filters=nr_filters, neighbourhood_size=1, dilation=dilation, bias=bias)
        self.bn = None
        self.relu = torch.nn.ReLU(inplace=False)

    def forward(self, lv, ls):
        ls.set_values(lv)
        if self.bn is None:
            self.bn = BatchNormLatticeModule(lv.shape[1], device=lv.device)
        lv, ls = self.bn(lv, ls)
        lv = self.relu(lv)
        ls.set_values(lv)
        lv_1, ls_1 = self.conv(lv, ls)
        ls_1.set_values(lv_1)
        return lv_1, ls_1


class CoarsenAct(torch.nn.Module):
    def __init__(self, in_channels, out_channels, device=None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.coarse = CoarsenLatticeModule(in_channels=in_channels, out_channels=out_channels, device=device)
        self.act = torch.nn.LeakyReLU(0.2)

    def forward(self, lv, ls, concat_connection=None):
        ls.set_values(lv)
        lv_1, ls_1 = self.coarse(lv, ls)
        lv_1 = self.act(lv_1)
        ls_1.set_values(lv_1)
        if concat_connection is not None:
            lv_1 = torch.cat((lv_1, concat_connection), 1)
            ls_1.set_values(lv_1)
        return lv_1, ls_1


class GnReluCoarsen(torch.nn.Module):
    def __init__(self, in_channels, out_channels, device=None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.coarse = CoarsenLatticeModule(in_channels=in_channels, out_channels=out_channels, device=device)
        self.norm = GroupNormLatticeModule(in_channels, device=device)
        self.relu = torch.nn.ReLU(inplace=False)

    def forward(self, lv, ls, concat_connection=None):
        ls.set_values(lv)
        lv, ls = self.norm.forward_relu(lv, ls)
        lv_1, ls_1 = self.coarse(lv, ls)
        ls_1.set_values(lv_1)
        if concat_connection is not None:
            lv_1 = torch.cat((lv_1, concat_connection), 1)
            ls_1.set_values(lv_1)
        return lv_1, ls_1


class FinefyAct(torch.nn.Module):
    def __init__(self, in_channels, out_channels, device=None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.fine = FinefyLatticeModule(in_channels=in_channels, out_channels=out_channels, device=device)
        self.act = torch.nn.LeakyReLU(0.2)

    def forward(self, lv_coarse, ls_coarse, ls_fine):
        ls_coarse.set_values(lv_coarse)
        lv_1, ls_1 = self.fine(lv_coarse, ls_coarse, ls_fine)
        lv_1 = self.act(lv_1)
        ls_1.set_values(lv_1)
        return lv_1, ls_1


class GnReluFinefy(torch.nn.Module):
    def __init__(self, in_channels, out_channels, device=None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.fine = FinefyLatticeModule(in_channels=in_channels, out_channels=out_channels, device=device)
        self.norm = GroupNormLatticeModule(in_channels, device=device)
        self.relu = torch.nn.ReLU(inplace=False)

    def forward(self, lv_coarse, ls_coarse, ls_fine):
        ls_coarse.set_values(lv_coarse)
        lv_coarse, ls_coarse = self.norm.forward_relu(lv_coarse, ls_coarse)
        lv_1, ls_1 = self.fine(lv_coarse, ls_coarse, ls_fine)
        ls_1.set_values(lv_1)
        return lv_1, ls_1


class ResnetBlock(torch.nn.Module):
    def __init__(self, in_channels, out_channels, dilations, biases, with_dropout, device=None):
        super().__init__()
        self.conv1 = GnReluConv(in_channels, out_channels, dilations[0], biases[0], with_dropout=False, device=device)
        self.conv2 = GnReluConv(in_channels, out_channels, dilations[1], biases[1], with_dropout=with_dropout, device=device)

    def forward(self, lv, ls):
        # lattice_modules.py:1255-1290: identity = lv; conv1; conv2; lv += identity.  The skip connection leaves through
        # the first GroupNorm node and re-enters in the epilogue of the second convolution: no add kernels either way.
        ls.set_values(lv)
        lv, identity, ls = self.conv1(lv, ls, split=True)
        lv, ls = self.conv2(lv, ls, residual=identity)
        ls.set_values(lv)
        return lv, ls


class BottleneckBlock(torch.nn.Module):
    """Pre-activation bottleneck: 1x1 down (/4), lattice conv, 1x1 up, skip (lattice_modules.py:1322-1358)."""

    def __init__(self, in_channels, out_channels, biases, device=None):
        super().__init__()
        self.downsample = 4
        mid = int(out_channels / self.downsample)
        self.contract = GnRelu1x1(in_channels=in_channels, out_channels=mid, bias=biases[0], device=device)
        self.conv = GnReluConv(in_channels=mid, out_channels=mid, dilation=1, bias=biases[1], with_dropout=False, device=device)
        self.expand = GnRelu1x1(in_channels=mid, out_channels=out_channels, bias=biases[2], device=device)

    def forward(self, lv, ls):
        ls.set_values(lv)
        lv, identity, ls = self.contract(lv, ls, split=True)
        lv, ls = self.conv(lv, ls)
        lv, ls = self.expand(lv, ls, residual=identity)
        ls.set_values(lv)
        return lv, ls


class SliceFastCUDALatticeModule(torch.nn.Module):
    """DeformSlice head: bottleneck -> gather -> learned barycentric offsets -> fused slice+classify
    (lattice_modules.py:465-567)."""

    def __init__(self, in_channels, nr_classes, dropout_prob, experiment, device=None):
        super().__init__()
        self.in_channels, self.nr_classes = in_channels, nr_classes
        self.bottleneck_size = 8
        self.stepdown = torch.nn.ModuleList([])
        self.linear_deltaW = None
        self.linear_clasify = None
        self.dropout_prob = dropout_prob
        if dropout_prob > 0.0:
            self.dropout = DropoutLattice(dropout_prob)
        self.experiment = experiment
        self.fused_delta_weights = True     # False: gather -> max -> affine -> Linear as separate torch ops (the reference's sequence)
        cur = in_channels
        for i in range(2):
            nr_out = int(in_channels / np.power(2, i))
            if nr_out < self.bottleneck_size:
                sys.exit("too many step-down layers: the bottleneck would expand instead of contract")
            self.stepdown.append(GnRelu1x1(cur, nr_out, False, device=device))
            cur = nr_out
        self.bottleneck = GnRelu1x1(cur, self.bottleneck_size, False, device=device)

    def forward(self, lv, ls, positions, splatting_indices, splatting_weights):
        ls.set_values(lv)
        assert self.in_channels == ls.val_dim(), \
            f"In channels doesn't match the val_dim of the lattice. In channels is {self.in_channels}, while val dim is {ls.val_dim()}"
        nr_positions = positions.shape[0]
        val_dim = lv.shape[1]
        lv_b, ls_b = lv, ls
        for layer in self.stepdown:
            lv_b, ls_b = layer(lv_b, ls_b)
        lv_b, ls_b = self.bottleneck(lv_b, ls_b)
        spv = ls.pos_dim() + 1
        per_vertex = self.bottleneck_size + 1
        fused = (self.fused_delta_weights and lv.is_cuda and self.bottleneck_size == 8 and spv in (4, 6) and self.experiment != "slice_no_deform"
                 and splatting_indices is not None and hasattr(ls, "m_hash_table"))
        if not fused:
            gathered = GatherLattice.apply(lv_b, ls_b, positions, splatting_indices, splatting_weights)
            per_vertex = int(gathered.shape[1] / spv)
        if self.linear_deltaW is None:
            self.linear_deltaW = torch.nn.Linear(per_vertex, 1, bias=True).to(lv.device)
            with torch.no_grad():
                torch.nn.init.kaiming_uniform_(self.linear_deltaW.weight, mode="fan_in", nonlinearity="tanh")
                self.linear_deltaW.weight *= 0.1   # start with offsets close to zero
                torch.nn.init.zeros_(self.linear_deltaW.bias)
            self.gamma = torch.nn.Parameter(torch.ones(per_vertex, device=lv.device))
            self.beta = torch.nn.Parameter(torch.zeros(per_vertex, device=lv.device))
        if fused:
            ls_b.set_values(lv_b)
            delta_weights = _DeltaWFn.apply(lv_b, splatting_indices, splatting_weights, self.gamma, self.beta, self.linear_deltaW.weight,
                                            self.linear_deltaW.bias, spv - 1)
        else:
            gathered = gathered.view(nr_positions, spv, per_vertex)
            max_vals, _ = gathered.max(1)
            gathered = gathered - (self.gamma * max_vals.unsqueeze(1) + self.beta)
            delta_weights = self.linear_deltaW(gathered).reshape(nr_positions, spv)
            if self.experiment == "slice_no_deform":
                delta_weights = delta_weights * 0
        if self.linear_clasify is None:
            self.linear_clasify = torch.nn.Linear(val_dim, self.nr_classes, bias=True).to(lv.device)
            leaky_relu_init_(self.linear_clasify.weight, val_dim + self.nr_classes, alpha=1.0)
            with torch.no_grad():
                self.linear_clasify.bias.zero_()
        if self.dropout_prob > 0.0:
            lv = self.dropout(lv)
        ls.set_values(lv)
        return SliceClassifyLattice.apply(lv, ls, positions, delta_weights, self.linear_clasify.weight,
                                          self.linear_clasify.bias, self.nr_classes, splatting_indices, splatting_weights)
