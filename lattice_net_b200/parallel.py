"""Scene-level data parallelism.  The reference is single-process /
single-GPU; each point cloud builds its own lattice, so the only natural shard is the scene.  One
process per GPU; the lattice is never split; the only exchange per step is ONE all-reduce of all
weight / bias gradients over NCCL (NVLink 5 / NVSwitch), done in place on a flat fp32 bucket that the
parameters' `.grad` tensors alias (no pack / unpack copies).

The model creates parameters lazily during its first forward (lattice_modules.py creates PointNet /
deltaW / classifier layers on first use), so the bucket is laid out after a warm-up forward --
the same reason the reference creates its optimizer late (ln_train.py:163-165).
"""
import os

import torch
import torch.distributed as dist


def init_distributed(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for a single process).
    Returns (rank, world_size, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local_rank


def scenes_for_rank(nr_scenes, rank, world):
    """Rank r takes scenes r, r+W, r+2W, ..."""
    return list(range(rank, nr_scenes, world))


class GradBucket:
    """All parameter gradients of a model in one flat fp32 buffer, for a single all-reduce per step.

    Gradients are produced by autograd as separate tensors (`.grad` is reset to None before each backward,
    so autograd hands its buffers over without an accumulation kernel per parameter); `pack()` gathers them
    into the flat buffer with one multi-tensor copy and re-points every `.grad` at its slice, so the
    optimizer reads the reduced values without an unpack pass."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, "no trainable parameters (run a warm-up forward first: they are created lazily)"
        dev = self.params[0].device
        # every slice starts on a 128-byte boundary: the lattice kernels write weight gradients into the slices with
        # 16-byte vector stores / reductions
        self.offsets, total = [], 0
        for p in self.params:
            self.offsets.append(total)
            total += (p.numel() + 31) // 32 * 32
        # one extra trailing element rides along with the gradients in the same collective (a per-step flag,
        # e.g. "this rank's cloud exceeded its vertex bound", graphed.py)
        self.flat_with_extra = torch.zeros(total + 1, dtype=torch.float32, device=dev)
        self.flat = self.flat_with_extra[:total]
        self.extra = self.flat_with_extra[total:]
        self.views = [self.flat[off:off + p.numel()].view_as(p) for p, off in zip(self.params, self.offsets)]
        self.nbytes = total * 4
        self._registered = False
        self._direct = False

    def zero(self):
        """Call before backward."""
        for p in self.params:
            p.grad = None

    def begin_direct_step(self):
        """Call before forward: clears the flat buffer with ONE memset and lets the backward passes of the lattice
        operators write weight gradients straight into their slices (lattice.grad_target); autograd then adopts those
        slices as `.grad` without a copy.  Pair with end_direct_step() after backward."""
        from . import lattice as _lattice
        if not self._registered:
            _lattice.register_grad_targets({p.data_ptr(): (self.flat, off, tuple(p.shape)) for p, off in zip(self.params, self.offsets)})
            self._registered = True
        self.flat_with_extra.zero_()
        for p in self.params:
            p.grad = None
        _lattice.set_grad_targets_active(True)
        self._direct = True

    def end_direct_step(self):
        from . import lattice as _lattice
        _lattice.set_grad_targets_active(False)

    def split_offset(self, first_late_param):
        """Flat offset of `first_late_param`: parameters from it to the end form the LATE chunk -- the layers nearest the loss,
        whose gradients are complete first in a backward pass -- so its all-reduce can overlap the rest of the backward."""
        for p, off in zip(self.params, self.offsets):
            if p is first_late_param:
                return off
        raise ValueError("parameter is not in this bucket")

    def pack(self, extra=None, first=0, last=None):
        """Call after backward: flat <- the grads that are not already there (one fused copy), .grad <- views of flat.
        first / last: restrict to the parameters whose slices start in [first, last) (flat offsets)."""
        dst, src = [], []
        last = self.flat.numel() if last is None else last
        for p, v, off in zip(self.params, self.views, self.offsets):
            if off < first or off >= last:
                continue
            if p.grad is None:
                if not self._direct:
                    v.zero_()                  # no memset cleared the buffer this step: drop the previous step's values
                continue
            if p.grad.data_ptr() != v.data_ptr():
                dst.append(v)
                src.append(p.grad)
        if dst:
            torch._foreach_copy_(dst, src)
        for p, v, off in zip(self.params, self.views, self.offsets):
            if first <= off < last:
                p.grad = v
        if extra is not None:
            self.extra.copy_(extra.reshape(1))

    def allreduce_mean(self, world):
        """One collective for the whole model; averages over ranks.  Single process: nothing to do."""
        if world > 1:
            self.pack()
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.mul_(1.0 / world)


def broadcast_parameters(model, src=0):
    """Make every rank start from rank `src`'s (lazily created) parameters, in one flat broadcast."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    params = [p for p in model.parameters()]
    flat = torch.cat([p.detach().reshape(-1) for p in params])
    dist.broadcast(flat, src)
    off = 0
    with torch.no_grad():
        for p in params:
            n = p.numel()
            p.copy_(flat[off:off + n].view_as(p))
            off += n
