"""Losses of the reference's training step: NLL on the log-softmax plus
the Lovasz-softmax surrogate of the mean IoU (Berman et al. 2018), which
/root/reference/latticenet_py/ln_train.py:156-158 combines 50/50.  ShapeNet-sized clouds
run it as one kernel (csrc/ln_train.cu); the batched torch formulation serves larger scans and CPU tensors."""
import torch


def lovasz_grad(gt_sorted):
    """Gradient of the Lovasz extension of the Jaccard loss w.r.t. sorted errors."""
    gts = gt_sorted.sum()
    intersection = gts - gt_sorted.cumsum(0)
    union = gts + (1.0 - gt_sorted).cumsum(0)
    jaccard = 1.0 - intersection / union
    if gt_sorted.numel() > 1:
        jaccard[1:] = jaccard[1:] - jaccard[:-1]
    return jaccard


def lovasz_softmax(probas, labels, ignore_index=None):
    """probas [P, C] (class probabilities), labels [P] (int64).  Mean over the classes present.
    All classes are handled by one batched sort / cumsum (the textbook version loops over classes).
    ignore_index: that CLASS is left out of the mean; its points stay in every other class's sorted error vector as
    negatives, exactly as in the reference (lovasz_loss.py:41-58 skips `c == ignore_index` in the class loop and never
    drops a point).  No boolean-mask indexing: static shapes, so this runs inside a CUDA-graph capture."""
    if probas.numel() == 0:
        return probas.sum() * 0.0
    nr_classes = probas.shape[1]
    # class-major [C, P] so that the sort and the scans run along the contiguous dimension; labels outside [0, C)
    # (e.g. a negative "unlabelled" marker) are foreground of no class
    in_range = (labels >= 0) & (labels < nr_classes)
    fg = torch.nn.functional.one_hot(labels.clamp(0, nr_classes - 1), nr_classes).to(probas.dtype) * in_range.unsqueeze(1).to(probas.dtype)
    fg = fg.t().contiguous()
    present = (fg.sum(1) > 0).to(probas.dtype)
    if ignore_index is not None and 0 <= ignore_index < nr_classes:
        present = present.clone()
        present[ignore_index] = 0.0
    errors_sorted, perm = torch.sort((fg - probas.t()).abs(), dim=1, descending=True)
    fg_sorted = fg.gather(1, perm)
    gts = fg_sorted.sum(1, keepdim=True)
    intersection = gts - fg_sorted.cumsum(1)
    union = gts + (1.0 - fg_sorted).cumsum(1)
    jaccard = 1.0 - intersection / union
    grad = torch.cat([jaccard[:, :1], jaccard[:, 1:] - jaccard[:, :-1]], 1)       # Lovasz extension gradient
    per_class = (errors_sorted * grad).sum(1)
    return (per_class * present).sum() / present.sum().clamp(min=1.0)


def lovasz_softmax_loop(probas, labels, ignore_index=None):
    """Per-class loop formulation (reference: latticenet_py/lattice/lovasz_loss.py:41-72), kept for tests."""
    losses = []
    for c in range(probas.shape[1]):
        if c == ignore_index:
            continue
        fg = (labels == c).to(probas.dtype)
        if fg.sum() == 0:
            continue
        errors = (fg - probas[:, c]).abs()
        errors_sorted, perm = torch.sort(errors, 0, descending=True)
        losses.append(torch.dot(errors_sorted, lovasz_grad(fg[perm])))
    return torch.stack(losses).mean()


class LovaszSoftmax(torch.nn.Module):
    def __init__(self, ignore_index=None):
        super().__init__()
        self.ignore_index = ignore_index

    def forward(self, logsoftmax, labels):
        return lovasz_softmax(logsoftmax.exp(), labels, self.ignore_index)


class _SegLossFn(torch.autograd.Function):
    """0.5 * Lovasz-softmax + 0.5 * NLL and its gradient w.r.t. the log-probabilities from ONE kernel (ln_seg_loss_fwd:
    one CTA per class sorts the errors in shared memory); the backward pass is a single scaling kernel."""

    @staticmethod
    def forward(ctx, logsoftmax, labels, ignore_index):
        from ._cabi import call, ptr, stream_ptr
        lp = logsoftmax.contiguous()
        n, c = lp.shape
        dev = lp.device
        grad_lov = torch.empty((n, c), dtype=torch.float32, device=dev)
        from . import lattice as _lattice
        scratch = _lattice._zeroed(1, 12, dev).view(-1)                        # [0:8) accumulators (zeroed), [8:12) result
        call("ln_seg_loss_fwd", ptr(lp), ptr(labels), n, c, int(ignore_index), ptr(grad_lov), ptr(scratch[:8]), ptr(scratch[8:]), stream_ptr(dev))
        ctx.save_for_backward(grad_lov, labels, scratch)
        ctx.ignore_index = int(ignore_index)
        return scratch[8]

    @staticmethod
    def backward(ctx, grad_loss):
        from ._cabi import call, ptr, stream_ptr
        grad_lov, labels, scratch = ctx.saved_tensors
        n, c = grad_lov.shape
        out = torch.empty_like(grad_lov)
        call("ln_seg_loss_bwd", ptr(grad_lov), ptr(labels), ptr(scratch[8:]), ptr(grad_loss.contiguous()), n, c, ctx.ignore_index, ptr(out),
             stream_ptr(out.device))
        return out, None, None


def segmentation_loss(logsoftmax, labels, ignore_index=-100):
    """0.5 * Lovasz-softmax + 0.5 * NLL, as in ln_train.py:156-158.  ShapeNet-sized clouds on the GPU run the fused
    kernel; larger scans (and CPU tensors) the batched torch formulation above."""
    if (logsoftmax.is_cuda and logsoftmax.dim() == 2 and logsoftmax.dtype == torch.float32 and labels.dtype == torch.int64
            and labels.is_cuda and labels.is_contiguous() and 1 <= logsoftmax.shape[0] <= _fused_loss_max_points()):
        return _SegLossFn.apply(logsoftmax, labels, ignore_index)
    nll = torch.nn.functional.nll_loss(logsoftmax, labels, ignore_index=ignore_index)
    lov = lovasz_softmax(logsoftmax.exp(), labels, ignore_index if ignore_index >= 0 else None)
    return 0.5 * lov + 0.5 * nll


_MAX_POINTS = None


def _fused_loss_max_points():
    global _MAX_POINTS
    if _MAX_POINTS is None:
        from ._cabi import load
        _MAX_POINTS = int(load().ln_seg_loss_max_points())
    return _MAX_POINTS
