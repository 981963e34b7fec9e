"""Scene-parallel plumbing on CPU: world_size-2 gloo run of the flat gradient bucket."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lattice_net_b200.parallel import GradBucket, broadcast_parameters, scenes_for_rank


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(rank)                      # ranks start different ...
    model = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    broadcast_parameters(model, 0)               # ... and must agree after the broadcast
    bucket = GradBucket(model.parameters())
    x = torch.full((5, 4), float(rank + 1))
    bucket.zero()
    model(x).sum().backward()
    local = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    bucket.allreduce_mean(world)
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    expected = sum(gathered) / world
    reduced = torch.cat([v.reshape(-1) for v in bucket.views])      # the flat buffer pads every slice to a 128-byte boundary
    ok = torch.allclose(reduced, expected) and all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in model.parameters())
    w0 = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    ws = [torch.zeros_like(w0) for _ in range(world)]
    dist.all_gather(ws, w0)
    ok = ok and torch.equal(ws[0], ws[1])
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_grad_bucket_allreduce_two_ranks():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_scene_sharding_is_a_partition():
    for world in (1, 2, 4, 8):
        seen = sorted(s for r in range(world) for s in scenes_for_rank(37, r, world))
        assert seen == list(range(37))


def test_bucket_pack_aliases_grads():
    model = torch.nn.Linear(3, 2)
    bucket = GradBucket(model.parameters())
    bucket.zero()
    assert all(p.grad is None for p in model.parameters())
    model(torch.ones(1, 3)).sum().backward()
    expect = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    bucket.pack()
    assert torch.equal(torch.cat([v.reshape(-1) for v in bucket.views]), expect)
    assert all(off % 32 == 0 for off in bucket.offsets) and float(bucket.flat.sum()) == float(expect.sum())     # padding stays zero
    lo, hi = bucket.flat.data_ptr(), bucket.flat.data_ptr() + bucket.nbytes
    assert all(lo <= p.grad.data_ptr() < hi for p in model.parameters())


def _two_chunk_worker(rank, world, port, out):
    """The protocol of graphed.GraphedTrainStep for world > 1, on CPU tensors: the LATE chunk (layers after the hook point,
    plus the per-step overflow flag) is packed and all-reduced from a backward hook while the earlier layers' gradients do
    not exist yet; the early chunk follows after the backward pass."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(7)
    early = torch.nn.Sequential(torch.nn.Linear(5, 6), torch.nn.Tanh(), torch.nn.Linear(6, 4))
    late = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Tanh(), torch.nn.Linear(3, 2))
    params = list(early.parameters()) + list(late.parameters())
    bucket = GradBucket(params)
    split = bucket.split_offset(next(iter(late.parameters())))
    x = torch.randn(9, 5, generator=torch.Generator().manual_seed(100 + rank))       # every rank its own scene
    overflow = torch.tensor(1.0 if rank == 1 else 0.0)                                  # only rank 1 exceeded its vertex bound
    state = {"early_grads_at_hook": None}

    def late_chunk_ready(_grad):
        state["early_grads_at_hook"] = [p.grad is None for p in early.parameters()]
        bucket.pack(extra=overflow, first=split)
        dist.all_reduce(bucket.flat_with_extra[split:], op=dist.ReduceOp.SUM)
        return None

    bucket.zero()
    bucket.flat_with_extra.zero_()
    h = early(x)
    h.register_hook(late_chunk_ready)
    late(h).pow(2).sum().backward()
    bucket.pack(last=split)
    dist.all_reduce(bucket.flat[:split], op=dist.ReduceOp.SUM)

    # the same gradients through one collective
    ref = GradBucket(params)
    for p in params:
        p.grad = None
    late(early(x)).pow(2).sum().backward()
    ref.pack(extra=overflow)
    dist.all_reduce(ref.flat_with_extra, op=dist.ReduceOp.SUM)

    ok = all(state["early_grads_at_hook"])                       # the hook really ran before the early layers' gradients existed
    ok = ok and 0 < split < bucket.flat.numel() and split % 32 == 0
    ok = ok and torch.equal(bucket.flat_with_extra, ref.flat_with_extra)
    ok = ok and float(bucket.extra) == 1.0                       # any rank's flag reaches every rank
    ok = ok and all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, ref.views))   # .grad re-pointed by the last pack
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_two_chunk_allreduce_equals_one_collective():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_two_chunk_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}
