"""Whole-step CUDA-graph execution of LatticeNet training (B200-first: the ShapeNet-sized step is
launch-bound -- ~550 kernels of a few microseconds each -- so the host, not the GPU, sets the pace of
the reference's loop, /root/reference/latticenet_py/ln_train.py:141-189).

A lattice is rebuilt for every cloud and its vertex count differs from cloud to cloud, which is what
normally rules out graph capture.  Here the lattice runs in *static-shape mode*
(`Lattice.set_vertex_bounds`): every level allocates a fixed number of rows, the actual vertex count
stays on the device (`nr_filled`), the kernels that need it (neighbour table, GroupNorm statistics)
read it there, and rows past it carry zeros / receive zero gradients.  Forward, backward, gradient
bucket and optimizer then replay as ONE graph launch per cloud.

A cloud whose lattice needs more vertices than a bound is detected on the device: the step's
optimizer update is skipped in-graph (fused AdamW `found_inf`), the event is counted, and the caller can
re-run that cloud through the ordinary (dynamic-shape) path.
"""
import torch
import torch.distributed as dist

from . import lattice as _lattice
from .lattice import Lattice


def estimate_vertex_bounds(capacity, sigmas, clouds, nr_levels, headroom=1.3, multiple=128):
    """Rows to reserve per lattice level: max vertex count over sample `clouds` (device tensors [N x d]),
    times `headroom`, rounded up to a multiple of `multiple` (128 = the M tile of the convolution kernel).
    Level l+1 is built like the model builds it: the raw points splatted at twice the sigma
    (Lattice::create_coarse_verts_naive, /root/reference/src/Lattice.cu:706-740)."""
    maxima = [0] * nr_levels
    for pos in clouds:
        lat = Lattice(capacity, sigmas)
        lat.begin_splat()
        lat.just_create_verts(pos, False)
        for lvl in range(nr_levels):
            maxima[lvl] = max(maxima[lvl], lat.nr_lattice_vertices())
            if lvl + 1 < nr_levels:
                lat = lat.create_coarse_verts_naive(pos)
    return [min(capacity, -(-int(m * headroom + 1) // multiple) * multiple) for m in maxima]


class GraphedTrainStep:
    """forward + loss + backward (+ gradient all-reduce) + optimizer step, captured once, replayed per cloud.

        step = GraphedTrainStep(model, lattice, optimizer, loss_fn, nr_points, val_dim, bounds, bucket, world)
        loss = step(positions, values, labels)          # device tensors or pinned host tensors

    `optimizer` must be created with capturable=True (fused AdamW supports it).  `bucket` is the
    parallel.GradBucket of the model (its flat buffer is what gets all-reduced when world > 1).
    """

    def __init__(self, model, lattice, optimizer, loss_fn, nr_points, pos_dim, val_dim, bounds, bucket, world=1,
                 warmup=3, capture_collective=False, example=None, overlap_allreduce=True):
        self.model, self.lattice, self.optimizer, self.loss_fn = model, lattice, optimizer, loss_fn
        self.bucket, self.world = bucket, world
        dev = bucket.flat.device
        self.device = dev
        lattice.set_vertex_bounds(bounds)
        self.bounds = list(bounds)
        self.pos = torch.zeros((nr_points, pos_dim), dtype=torch.float32, device=dev)
        self.vals = torch.zeros((nr_points, val_dim), dtype=torch.float32, device=dev)
        self.labels = torch.zeros((nr_points,), dtype=torch.int64, device=dev)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self.found_inf = torch.zeros((), dtype=torch.float32, device=dev)      # 1.0 = a bound was exceeded: skip the update
        self.overflow_steps = torch.zeros((), dtype=torch.float32, device=dev)  # running count of such steps
        self.nv_actual = torch.zeros((len(self.bounds),), dtype=torch.int32, device=dev)
        if example is not None:                 # a representative cloud for the warm-up / capture passes
            self.pos.copy_(example[0])
            self.vals.copy_(example[1])
            self.labels.copy_(example[2])
        optimizer.found_inf = self.found_inf      # torch fused AdamW (capturable) and optim.FlatAdamW both honour it on the device
        self.flat_optimizer = hasattr(optimizer, "flat_params")
        self.capture_collective = capture_collective and world > 1
        self.graphs = []
        self.launches_per_step = 0
        # zeroed outputs of the split-K convolutions: one buffer, one memset per step (sized by a dry run in _capture)
        self.arena = _lattice.ZeroArena(0, None)
        # world > 1: the gradient all-reduce runs in two chunks.  The LATE chunk (decoder + slice head, ~80 % of the bytes,
        # complete when the backward pass reaches the bottleneck) goes out on a side stream under the encoder's backward;
        # the early chunk follows at the end.  Both are captured inside the step graph.
        self.split = None
        self.side = None
        if self.capture_collective and overlap_allreduce and hasattr(model, "finefy_list") and hasattr(model, "grad_sync_hook"):
            first_late = next(iter(model.finefy_list.parameters()), None)
            if first_late is not None:
                self.split = bucket.split_offset(first_late)
                self.side = torch.cuda.Stream(device=dev)
                model.grad_sync_hook = self._late_chunk_ready
        self._capture(warmup)

    # -- the three phases of a step; `_allreduce` is the only part that may have to stay outside a graph
    def _levels_status(self):
        import ctypes
        from ._cabi import call, ptr, stream_ptr
        sts = [l.m_hash_table.structure for l in self.model.last_level_lattices]
        call("ln_levels_status", (ctypes.c_void_p * len(sts))(*[s.nr_filled.data_ptr() for s in sts]),
             (ctypes.c_void_p * len(sts))(*[s.status.data_ptr() for s in sts]), len(sts), ptr(self.nv_actual), ptr(self.found_inf),
             stream_ptr(self.device))

    def _late_chunk_ready(self, _grad):
        """Backward-pass hook at the bottleneck (LNN.forward): the gradients of every later layer are final.  Collect the few
        that are not in the bucket yet, append the overflow flag, and all-reduce that chunk on the side stream."""
        _lattice.join_deferred_wgrads(self.device)           # the late layers' weight gradients trail on the side stream
        self.bucket.pack(extra=self.found_inf, first=self.split)
        cur = torch.cuda.current_stream(self.device)
        self.side.wait_stream(cur)
        with torch.cuda.stream(self.side):
            dist.all_reduce(self.bucket.flat_with_extra[self.split:], op=dist.ReduceOp.SUM)
        self._late_sent = True
        return None

    def _forward_backward(self):
        self.arena.reset()                       # one memset for every output a split-K convolution accumulates into
        self.bucket.begin_direct_step()          # one memset for every weight gradient; .grad <- None
        self._late_sent = False
        prev = _lattice.set_zero_arena(self.arena)
        prev_defer = _lattice.set_defer_wgrad_join(True)     # weight gradients trail the data-gradient chain; joined below
        try:
            logsoftmax, _ = self.model(self.lattice, self.pos, self.vals)
            self._levels_status()                # the lattice pyramid is complete after the forward pass: overflow flag, vertex counts
            loss = self.loss_fn(logsoftmax, self.labels)
            loss.backward()
        finally:
            _lattice.set_defer_wgrad_join(prev_defer)
            _lattice.join_deferred_wgrads(self.device)
            _lattice.set_zero_arena(prev)
            self.bucket.end_direct_step()
        self.loss.copy_(loss.detach())
        if self.world > 1 or self.flat_optimizer:
            # gradients that did not land in the bucket by themselves (a handful: biases, weight-norm gains) are copied in
            if self._late_sent:
                self.bucket.pack(last=self.split)
            else:
                self.bucket.pack(extra=self.found_inf)

    def _allreduce(self):
        if self.world > 1:
            if self._late_sent:
                dist.all_reduce(self.bucket.flat[:self.split], op=dist.ReduceOp.SUM)
                torch.cuda.current_stream(self.device).wait_stream(self.side)
            else:
                dist.all_reduce(self.bucket.flat_with_extra, op=dist.ReduceOp.SUM)

    def _update(self):
        if self.world > 1:
            self.found_inf.copy_((self.bucket.extra > 0).to(torch.float32).reshape(()))   # any rank overflowed -> all skip
        self.overflow_steps.add_(self.found_inf)
        if self.flat_optimizer:
            self.optimizer.step(grad_scale=1.0 / self.world)         # the 1/W of the sum all-reduce rides in the update kernel
        else:
            if self.world > 1:
                self.bucket.flat.mul_(1.0 / self.world)
            self.optimizer.step()

    def _eager_step(self):
        self._forward_backward()
        self._allreduce()
        self._update()

    def _snapshot(self):
        if self.flat_optimizer:
            o = self.optimizer
            return [t.clone() for t in (o.flat_params, o.exp_avg, o.exp_avg_sq, o.max_exp_avg_sq, o.state)]
        params = [p for g in self.optimizer.param_groups for p in g["params"]]
        saved = []
        for p in params:
            st = self.optimizer.state.get(p, None)
            saved.append((p, p.detach().clone(), None if not st else {k: v.clone() for k, v in st.items() if torch.is_tensor(v)}))
        return saved

    def _restore(self, saved):
        # in place: the graph has captured the addresses of the parameters and of the optimizer state
        if self.flat_optimizer:
            o = self.optimizer
            with torch.no_grad():
                for dst, src in zip((o.flat_params, o.exp_avg, o.exp_avg_sq, o.max_exp_avg_sq, o.state), saved):
                    dst.copy_(src)
            return
        with torch.no_grad():
            for p, value, st in saved:
                p.copy_(value)
                for k, v in self.optimizer.state.get(p, {}).items():
                    if torch.is_tensor(v):
                        v.copy_(st[k]) if st is not None else v.zero_()

    def _capture(self, warmup):
        from . import _cabi
        saved = self._snapshot()                # the warm-up passes are real optimizer steps: undo them afterwards
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            self._eager_step()                  # dry run of the arena: counts the floats the step asks for
            self.arena = _lattice.ZeroArena(self.arena.requested + 1024, self.device)
            for _ in range(max(warmup, 1)):     # lazy initialisations (cuBLAS handles, workspaces, optimizer state)
                self._eager_step()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.overflow_steps.zero_()
        _cabi.reset_launch_count()
        if self.world == 1 or self.capture_collective:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._eager_step()
            self.graphs = [g]
            self._between = None
        else:
            ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(ga):
                self._forward_backward()
            with torch.cuda.graph(gb, pool=ga.pool()):
                self._update()
            self.graphs = [ga, gb]
            self._between = self._allreduce
        self.launches_per_step = _cabi.launch_count()      # kernels of this library inside one replay
        self._restore(saved)
        self.overflow_steps.zero_()

    def __call__(self, positions, values, labels):
        self.pos.copy_(positions, non_blocking=True)
        self.vals.copy_(values, non_blocking=True)
        self.labels.copy_(labels, non_blocking=True)
        self.graphs[0].replay()
        if self._between is not None:
            self._between()
            self.graphs[1].replay()
        return self.loss

    def overflowed_steps(self):
        """Number of replayed steps whose cloud exceeded a vertex bound (blocking read)."""
        return int(self.overflow_steps.item())

    def last_vertex_counts(self):
        return [int(v) for v in self.nv_actual.tolist()]
