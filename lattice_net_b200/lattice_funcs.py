"""Autograd Functions of the lattice operators -- same names, forward signatures and backward
return arities as /root/reference/latticenet_py/lattice/lattice_funcs.py:30-603, on top of the
B200 `Lattice` handle.

Convolution-type Functions (ConvIm2Row / Coarsen / Finefy) differ from the reference in mechanism
only: the reference's backward re-materialises im2row and calls `mm` twice (lattice_funcs.py:294-313);
here the weight gradient and the data gradient are one implicit-GEMM kernel each over the cached
neighbour table.
"""
import sys

import torch
from torch.autograd import Function

from . import lattice as _lattice_mod
from .lattice_wrapper import LatticeWrapper


def _conv_backward(query, neighbours, neighbour_values, filter_bank, grad_out, dilation, val_dim, need_input_grad=True):
    """Shared backward of out = conv(query <- neighbours):
    grad_filter = im2row(neighbours)^T . grad_out;  grad_neighbour_values = flipped conv of grad_out
    evaluated at the neighbour lattice's vertices with the re-laid-out filter (lattice_funcs.py:298-313)."""
    filter_extent = int(filter_bank.shape[0] // val_dim)
    grad_out = grad_out.contiguous()
    neighbours.set_values(neighbour_values)
    if hasattr(query, "conv_backward"):     # both gradients from one call into the CUDA library
        # a bank that is a leaf parameter gets its gradient written straight into its gradient-bucket slice (when one is active)
        grad_param = filter_bank if (filter_bank.is_leaf and filter_bank.requires_grad) else None
        grad_values, grad_filter = query.conv_backward(neighbours, grad_out, filter_bank, dilation, need_input_grad, grad_param)
        query.set_values(grad_out)          # the reference leaves the query handle holding the incoming gradient
        return grad_values, grad_filter
    grad_filter = query.conv_weight_grad(neighbours, grad_out, filter_extent, dilation)
    query.set_values(grad_out)
    if getattr(neighbours, "SUPPORTS_TRANSPOSED_FILTER", False):
        # the kernel reads the forward bank transposed in place: no re-laid-out copy of the filter
        grad_lattice = neighbours.convolve_im2row_standalone(filter_bank, dilation, query, True, transposed_filter=True)
    else:
        filter_bw = query.filter_for_data_grad(filter_bank, filter_extent, val_dim)
        grad_lattice = neighbours.convolve_im2row_standalone(filter_bw, dilation, query, True)
    return grad_lattice.values(), grad_filter


class SplatLattice(Function):
    @staticmethod
    def forward(ctx, lattice, positions, values):
        lattice.begin_splat()
        indices, weights = lattice.splat_standalone(positions, values)
        return lattice.values(), LatticeWrapper.wrap(lattice), indices, weights

    @staticmethod
    def backward(ctx, grad_lattice_values, grad_lattice_structure, grad_indices=None, grad_weights=None):
        return None, None, None   # no gradient flows through a splat (lattice_funcs.py:41-43)


class DistributeLattice(Function):
    @staticmethod
    def forward(ctx, lattice, positions, values, reset_hashmap=True):
        lattice.begin_splat(reset_hashmap)
        distributed_lattice, distributed, indices, weights = lattice.distribute(positions, values, reset_hashmap)
        ctx.save_for_backward(indices, weights)
        ctx.pos_dim = lattice.pos_dim()
        ctx.val_dim = lattice.val_dim()
        ctx.nr_positions = positions.shape[0]
        return LatticeWrapper.wrap(distributed_lattice), distributed, indices, weights

    @staticmethod
    def backward(ctx, grad_wrap, grad_distributed, grad_indices, grad_weights):
        d, v, n = ctx.pos_dim, ctx.val_dim, ctx.nr_positions
        # each point's value was copied to its d+1 simplex vertices: sum their gradients
        grad_values = grad_distributed[:, d:d + v].reshape(n, d + 1, v).sum(dim=1)
        return None, None, grad_values, None


class ExpandLattice(Function):
    @staticmethod
    def forward(ctx, lattice_values, lattice_structure, positions, point_multiplier, noise_stddev, expand_values):
        lattice_structure.set_values(lattice_values)
        expanded = lattice_structure.expand(positions, point_multiplier, noise_stddev, expand_values)
        ctx.nr_values_original_lattice = lattice_structure.nr_lattice_vertices()
        return expanded.values(), LatticeWrapper.wrap(expanded)

    @staticmethod
    def backward(ctx, grad_lattice_values, grad_lattice_structure):
        return grad_lattice_values[0:ctx.nr_values_original_lattice, :], None, None, None, None, None


class Im2RowIndicesLattice(Function):
    @staticmethod
    def forward(ctx, lattice_values, lattice, filter_extent, dilation, nr_filters):
        lattice.set_values(lattice_values)
        ctx.lattice, ctx.filter_extent, ctx.dilation, ctx.nr_filters = lattice, filter_extent, dilation, nr_filters
        return lattice.im2rowindices(lattice, filter_extent, dilation, False)

    @staticmethod
    def backward(ctx, grad_lattice_rowified):
        lattice = ctx.lattice
        grad_values = lattice.row2im(grad_lattice_rowified.contiguous(), ctx.dilation, ctx.filter_extent, ctx.nr_filters, lattice)
        ctx.lattice = None
        return grad_values, None, None, None, None


class Im2RowLattice(Function):
    @staticmethod
    def forward(ctx, lattice_values, lattice, filter_extent, dilation, nr_filters):
        lattice.set_values(lattice_values)
        ctx.lattice, ctx.filter_extent, ctx.dilation, ctx.nr_filters = lattice, filter_extent, dilation, nr_filters
        ctx.val_dim = lattice.val_dim()
        return lattice.im2row(lattice, filter_extent, dilation, False)

    @staticmethod
    def backward(ctx, grad_lattice_rowified):
        lattice = ctx.lattice
        if lattice.val_dim() != ctx.val_dim:
            # the handle's values were replaced since the forward; restore the width row2im expects
            lattice.m_hash_table.m_values_tensor = grad_lattice_rowified.new_zeros((lattice.nr_lattice_vertices(), ctx.val_dim))
        grad_values = lattice.row2im(grad_lattice_rowified.contiguous(), ctx.dilation, ctx.filter_extent, ctx.nr_filters, lattice)
        ctx.lattice = None
        return grad_values, None, None, None, None


class ConvIm2RowLattice(Function):
    """forward(lattice_values, lattice, filter_bank, dilation) as in the reference (lattice_funcs.py:250-320); the two
    optional trailing inputs are extensions: `bias` [nr_filters] and `residual` [nv x nr_filters] are added in the
    convolution kernel's epilogue (the reference adds them with separate torch ops, lattice_modules.py:244-246, 1290)."""

    @staticmethod
    def forward(ctx, lattice_values, lattice, filter_bank, dilation, bias=None, residual=None):
        lattice.set_values(lattice_values)
        if bias is None and residual is None:
            convolved = lattice.convolve_im2row_standalone(filter_bank, dilation, lattice, False)
        else:
            convolved = lattice.convolve_im2row_standalone(filter_bank, dilation, lattice, False, bias=bias, residual=residual)
        ctx.save_for_backward(filter_bank, lattice_values)
        ctx.lattice, ctx.dilation, ctx.val_dim = lattice, dilation, lattice.val_dim()
        ctx.has_bias, ctx.has_residual = bias is not None, residual is not None
        return convolved.values(), LatticeWrapper.wrap(convolved)

    @staticmethod
    def backward(ctx, grad_lattice_values, grad_lattice_structure):
        filter_bank, lattice_values = ctx.saved_tensors
        lattice = ctx.lattice
        # same lattice on both sides: query a private alias so set_values on one role does not clobber the other
        query = lattice.clone_lattice()
        grad_values, grad_filter = _conv_backward(query, lattice, lattice_values, filter_bank, grad_lattice_values, ctx.dilation, ctx.val_dim,
                                                  ctx.needs_input_grad[0])
        ctx.lattice = None
        grad_bias = grad_lattice_values.sum(0) if (ctx.has_bias and ctx.needs_input_grad[4]) else None
        grad_residual = grad_lattice_values if (ctx.has_residual and ctx.needs_input_grad[5]) else None
        return grad_values, None, grad_filter, None, grad_bias, grad_residual


class CoarsenLattice(Function):
    @staticmethod
    def forward(ctx, lattice_fine_values, lattice_fine_structure, filter_bank, coarsened_lattice=None):
        lattice_fine_structure.set_values(lattice_fine_values)
        if coarsened_lattice is None:
            coarsened_lattice = lattice_fine_structure.create_coarse_verts_naive(lattice_fine_structure.positions())
        dilation = 1
        convolved = coarsened_lattice.convolve_im2row_standalone(filter_bank, dilation, lattice_fine_structure, False)
        ctx.save_for_backward(filter_bank, lattice_fine_values)
        ctx.coarsened_lattice, ctx.lattice_fine_structure = coarsened_lattice, lattice_fine_structure
        ctx.dilation, ctx.val_dim = dilation, lattice_fine_structure.val_dim()
        return convolved.values(), LatticeWrapper.wrap(convolved)

    @staticmethod
    def backward(ctx, grad_lattice_values, grad_lattice_structure):
        filter_bank, lattice_fine_values = ctx.saved_tensors
        grad_fine, grad_filter = _conv_backward(ctx.coarsened_lattice, ctx.lattice_fine_structure, lattice_fine_values,
                                                filter_bank, grad_lattice_values, ctx.dilation, ctx.val_dim)
        ctx.coarsened_lattice = ctx.lattice_fine_structure = None
        return grad_fine, None, grad_filter, None


class FinefyLattice(Function):
    @staticmethod
    def forward(ctx, lattice_coarse_values, lattice_coarse_structure, lattice_fine_structure, filter_bank):
        lattice_coarse_structure.set_values(lattice_coarse_values)
        dilation = 1
        convolved = lattice_fine_structure.convolve_im2row_standalone(filter_bank, dilation, lattice_coarse_structure, False)
        ctx.save_for_backward(filter_bank, lattice_coarse_values)
        ctx.lattice_fine_structure, ctx.lattice_coarse_structure = convolved, lattice_coarse_structure
        ctx.dilation, ctx.val_dim = dilation, lattice_coarse_structure.val_dim()
        return convolved.values(), LatticeWrapper.wrap(convolved)

    @staticmethod
    def backward(ctx, grad_lattice_values, grad_lattice_structure):
        filter_bank, lattice_coarse_values = ctx.saved_tensors
        grad_coarse, grad_filter = _conv_backward(ctx.lattice_fine_structure, ctx.lattice_coarse_structure, lattice_coarse_values,
                                                  filter_bank, grad_lattice_values, ctx.dilation, ctx.val_dim)
        ctx.lattice_fine_structure = ctx.lattice_coarse_structure = None
        return grad_coarse, None, None, grad_filter


class SliceLattice(Function):
    @staticmethod
    def forward(ctx, lattice_values, lattice_structure, positions, splatting_indices=None, splatting_weights=None):
        lattice_structure.set_values(lattice_values)
        if splatting_indices is None and splatting_weights is None:
            sliced, splatting_indices, splatting_weights = lattice_structure.slice_standalone_no_precomputation(positions)
        else:
            sliced = lattice_structure.slice_standalone_with_precomputation(positions, splatting_indices, splatting_weights)
        ctx.save_for_backward(positions, splatting_indices, splatting_weights)
        ctx.lattice_structure = lattice_structure
        return sliced

    @staticmethod
    def backward(ctx, grad_sliced_values):
        positions, splatting_indices, splatting_weights = ctx.saved_tensors
        lattice_structure = ctx.lattice_structure
        if lattice_structure.val_dim() != grad_sliced_values.shape[1]:
            sys.exit("the values stored in the lattice do not have the dimension of the gradient")   # lattice_funcs.py:504-505
        lattice_structure.slice_backwards_standalone_with_precomputation_no_homogeneous(
            positions, grad_sliced_values.contiguous(), splatting_indices, splatting_weights)
        grad_values = lattice_structure.values()
        ctx.lattice_structure = None
        return grad_values, None, None, None, None


class SliceClassifyLattice(Function):
    @staticmethod
    def forward(ctx, lattice_values, lattice_structure, positions, delta_weights, linear_clasify_weight,
                linear_clasify_bias, nr_classes, splatting_indices, splatting_weights):
        lattice_structure.set_values(lattice_values)
        if lattice_values.is_cuda and any(ctx.needs_input_grad):
            # the limits of the backward kernels (ln_slice_classify_bwd), checked here so that an unsupported head fails
            # before training starts rather than in the first backward()
            v, nc = int(lattice_values.shape[1]), int(nr_classes)
            if v > 256 or nc * v > 6144 or nc > 256:
                raise RuntimeError(f"slice_classify backward is built for val_dim <= 256, nr_classes <= 256 and nr_classes * val_dim <= 6144 "
                                   f"(got val_dim {v}, nr_classes {nc}); run the head under torch.no_grad() or use SliceLattice + a Linear layer")
        logits = lattice_structure.slice_classify_with_precomputation(
            positions, delta_weights, linear_clasify_weight, linear_clasify_bias, nr_classes, splatting_indices, splatting_weights)
        ctx.save_for_backward(positions, lattice_values, delta_weights, linear_clasify_weight, linear_clasify_bias,
                              splatting_indices, splatting_weights)
        ctx.lattice_structure, ctx.nr_classes = lattice_structure, nr_classes
        return logits

    @staticmethod
    def backward(ctx, grad_class_logits):
        positions, initial_values, delta_weights, cls_w, cls_b, splatting_indices, splatting_weights = ctx.saved_tensors
        lattice = ctx.lattice_structure
        zeroed = getattr(_lattice_mod, "_zeroed", None)
        if initial_values.is_cuda and zeroed is not None:
            # one memset per step clears the arena these come from (graphed step); the classifier gradients go straight
            # into their gradient-bucket slices when a zeroed bucket is active
            grad_lattice_values = zeroed(initial_values.shape[0], initial_values.shape[1], initial_values.device)
            grad_delta_weights = zeroed(delta_weights.shape[0], delta_weights.shape[1], delta_weights.device)
            grad_cls_w = _lattice_mod.grad_target(cls_w) if cls_w.is_leaf else None
            grad_cls_b = _lattice_mod.grad_target(cls_b) if cls_b.is_leaf else None
            if grad_cls_w is None or grad_cls_b is None:
                grad_cls_w, grad_cls_b = torch.zeros_like(cls_w), torch.zeros_like(cls_b)
        else:
            grad_lattice_values = torch.zeros_like(initial_values)
            grad_delta_weights = torch.zeros_like(delta_weights)
            grad_cls_w = torch.zeros_like(cls_w)
            grad_cls_b = torch.zeros_like(cls_b)
        lattice.slice_classify_backwards_with_precomputation(
            grad_class_logits.contiguous(), positions, initial_values, delta_weights, cls_w, cls_b, ctx.nr_classes,
            grad_lattice_values, grad_delta_weights, grad_cls_w, grad_cls_b, splatting_indices, splatting_weights)
        ctx.lattice_structure = None
        return grad_lattice_values, None, None, grad_delta_weights, grad_cls_w, grad_cls_b, None, None, None


class GatherLattice(Function):
    @staticmethod
    def forward(ctx, lattice_values, lattice_structure, positions, splatting_indices, splatting_weights):
        lattice_structure.set_values(lattice_values)
        gathered = lattice_structure.gather_standalone_with_precomputation(positions, splatting_indices, splatting_weights)
        ctx.save_for_backward(positions, splatting_indices, splatting_weights)
        ctx.lattice_structure = lattice_structure
        return gathered

    @staticmethod
    def backward(ctx, grad_sliced_values):
        positions, splatting_indices, splatting_weights = ctx.saved_tensors
        lattice = ctx.lattice_structure
        lattice.gather_backwards_standalone_with_precomputation(positions, grad_sliced_values.contiguous(), splatting_indices, splatting_weights)
        grad_values = lattice.values()
        ctx.lattice_structure = None
        return grad_values, None, None, None, None
