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
