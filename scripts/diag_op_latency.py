"""In-graph latency of dependent chains of the step's building blocks at ShapeNet sizes (what one more / one fewer kernel of
each kind costs inside the captured step).  usage (GPU box): python scripts/diag_op_latency.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def chain_time(fn, reps=20, replays=20):
    dev = torch.device("cuda", 0)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(2):
            fn()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * replays)


def main():
    from lattice_net_b200 import Lattice, _cabi, lattice as L
    from lattice_net_b200.lattice_modules import _GroupNormReLU, linear
    dev = torch.device("cuda", 0)
    pos = torch.from_numpy(bench.synthetic_cloud(7)[0]).to(dev)
    lat = Lattice(bench.CAPACITY, [(bench.SIGMA, 3)])
    lat.set_vertex_bounds([1408, 384, 128, 128])
    lat.begin_splat()
    lat.just_create_verts(pos, False)
    levels = [lat]
    for _ in range(3):
        levels.append(levels[-1].create_coarse_verts_naive(pos))
    arena = L.ZeroArena(64 << 20, dev)
    L.set_zero_arena(arena)
    res = {}
    for pdl in (1, 0):
        _cabi.load().ln_set_programmatic_launch(pdl)
        for lvl, cin, cout in [(0, 32, 32), (0, 128, 128), (1, 64, 64), (1, 192, 192), (2, 32, 32), (3, 64, 64)]:
            h = levels[lvl].clone_lattice()
            nv = h.nr_lattice_vertices()
            state = {"x": torch.randn((nv, cin), device=dev)}
            fb = torch.randn((9 * cin, cout), device=dev) * 0.05
            L.prepare_filters([(fb, 9, cin, cout, False), (fb, 9, cout, cin, True)])
            h.set_values(state["x"])

            def conv():
                arena.off = 0
                h.convolve_im2row_standalone(fb, 1, h, False)

            res[f"conv {cin}->{cout} rows={nv} pdl={pdl}"] = chain_time(conv)
            g = torch.randn((nv, cout), device=dev)
            q = h.clone_lattice()

            def bwd():
                arena.off = 0
                q.conv_backward(h, g, fb, 1)

            res[f"conv_bwd(dgrad||wgrad) {cin}->{cout} rows={nv} pdl={pdl}"] = chain_time(bwd)
        for nv, c in [(1408, 32), (1408, 128), (384, 64), (128, 128), (128, 256)]:
            x = torch.randn((nv, c), device=dev)
            gam, bet = torch.ones(c, device=dev), torch.zeros(c, device=dev)
            res[f"group_norm fwd rows={nv} C={c} pdl={pdl}"] = chain_time(lambda: _GroupNormReLU.apply(x, gam, bet, 32, 1e-5, True, None))
        for m, k, n in [(128, 128, 32), (128, 32, 128), (1408, 128, 128)]:
            x = torch.randn((m, k), device=dev)
            w = torch.randn((n, k), device=dev) * 0.05
            L.prepare_filters([(w, 1, k, n, True)])

            def lin():
                arena.off = 0
                linear(x, w)

            res[f"linear {m}x{k}->{n} pdl={pdl}"] = chain_time(lin)
        st = levels[0].m_hash_table.structure
        res[f"table_clear (tiny kernel) pdl={pdl}"] = chain_time(lambda: st.clear())
        t = torch.zeros(1000, device=dev)
        res[f"torch add_ (tiny kernel) pdl={pdl}"] = chain_time(lambda: t.add_(1.0))
    for k, v in res.items():
        print(f"{v:8.2f} us  {k}")


if __name__ == "__main__":
    main()
