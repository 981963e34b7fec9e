"""AdamW with amsgrad as ONE kernel over flat buffers.

The reference trains with `torch.optim.AdamW(model.parameters(), lr, weight_decay, amsgrad=True)`
(/root/reference/latticenet_py/ln_train.py:163-165).  torch's fused implementation walks the 154 parameter tensors of
LatticeNet in six multi-tensor launches; here the parameters are re-homed into one flat buffer (each `p.data` becomes
a view of it), the gradients already live in `parallel.GradBucket.flat`, and the update -- same arithmetic, same order --
is a single launch of `ln_adamw_amsgrad` that also applies the 1/world_size of the gradient all-reduce and honours the
device-side "skip this step" flag of the graphed step.
"""
import torch

from ._cabi import call, ptr, stream_ptr


class FlatAdamW:
    """opt = FlatAdamW(bucket, lr=1e-3, weight_decay=3e-4);  ...backward...;  bucket.pack();  opt.step()

    `bucket` is the parallel.GradBucket of the model: its parameter order defines the flat layout, and after
    `bucket.pack()` its flat buffer holds every gradient."""

    def __init__(self, bucket, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        from . import lattice as _lattice
        self.bucket = bucket
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        params = bucket.params
        dev = params[0].device
        n = bucket.flat.numel()             # the bucket's layout (slices padded to 128-byte boundaries) is the layout of everything here
        self.n = n
        self.flat_params = torch.zeros((n,), dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, off in zip(params, bucket.offsets):
                view = self.flat_params[off:off + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view               # the parameter now lives inside the flat buffer
        _lattice.invalidate_prepared_filters()      # prepared slabs and gradient targets are keyed by the old addresses
        bucket._registered = False
        self.exp_avg = torch.zeros_like(self.flat_params)
        self.exp_avg_sq = torch.zeros_like(self.flat_params)
        self.max_exp_avg_sq = torch.zeros_like(self.flat_params)
        self.state = torch.zeros((2,), dtype=torch.float32, device=dev)     # [step count, scratch]
        self.found_inf = None               # optional device float: non-zero = skip the update (set by GraphedTrainStep)
        self.param_groups = [{"params": params, "lr": self.lr}]               # enough of the torch.optim surface for schedulers / logging

    def zero_grad(self, set_to_none=True):
        self.bucket.zero()

    def step(self, grad_scale=1.0):
        lr = float(self.param_groups[0]["lr"])
        call("ln_adamw_amsgrad", ptr(self.flat_params), ptr(self.bucket.flat), ptr(self.exp_avg), ptr(self.exp_avg_sq),
             ptr(self.max_exp_avg_sq), self.n, lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, float(grad_scale),
             ptr(self.state), ptr(self.found_inf), stream_ptr(self.flat_params.device))

    def steps_taken(self):
        return int(self.state[0].item())

    def state_dict(self):
        return {"exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "max_exp_avg_sq": self.max_exp_avg_sq, "state": self.state,
                "lr": self.param_groups[0]["lr"], "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay}

    def load_state_dict(self, sd):
        with torch.no_grad():
            for k in ("exp_avg", "exp_avg_sq", "max_exp_avg_sq", "state"):
                getattr(self, k).copy_(sd[k])
        self.param_groups[0]["lr"] = sd["lr"]
