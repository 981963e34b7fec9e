"""TEST INFRASTRUCTURE (oracle) -- torch-CPU re-expression of the whole LatticeNet step.

The reference has no CPU implementation (README.md:20).  This port restates the reference's semantics
with plain differentiable torch-CPU ops -- lattice structure from the C oracle + np.unique, splat /
pooling as index_add / scatter_reduce, convolution as neighbour-table gather + mm, slice as gather +
weighted sum -- so autograd provides the backward pass.  It is used (1) as the `cpu_baseline` of
bench.py (kind "port", timed on the box's host cores) and (2) by the tests as a model-level parity check
of the CUDA path (same parameters, same cloud -> same logits and gradients).

Module / parameter names equal those of lattice_net_b200.models.LNN so a state_dict loads one to one.
"""
import os
import time

import numpy as np
import torch
import torch.nn.functional as F_

from . import lattice_oracle as lo


# --------------------------------------------------------------------------------------------------
class Level:
    """Structure of one lattice level on the CPU."""

    def __init__(self, keys):
        self.keys = np.asarray(keys, np.int32)
        self.nv = len(self.keys)
        self._tables = {}

    def table(self, other, lvl_diff, dilation=1):
        key = (id(other), lvl_diff, dilation)
        if key not in self._tables:
            t = lo.neighbour_table(self.keys, other.keys, lvl_diff, dilation)
            t = np.where(t < 0, other.nv, t)                 # absent -> index of the zero padding row
            self._tables[key] = torch.from_numpy(t.astype(np.int64))
        return self._tables[key]


def lattice_conv(x, table, weight, flip=False):
    """rows = gather(x, table) -> [nv_q, F*C] ; rows @ weight   (Lattice.cu:424-474)"""
    nvq, Fe = table.shape
    xp = torch.cat([x, x.new_zeros(1, x.shape[1])], 0)
    if flip:
        perm = [s ^ 1 for s in range(Fe - 1)] + [Fe - 1]
        table = table[:, perm]
    rows = xp[table.reshape(-1)].reshape(nvq, Fe * x.shape[1])
    return rows.mm(weight)


def group_norm(x, gn_w, gn_b):
    c = x.shape[1]
    groups = 32 if c % 32 == 0 else c // 2
    return F_.group_norm(x.t().unsqueeze(0), groups, gn_w, gn_b).squeeze(0).t()


def wn(v, g):
    return v * (g / v.norm())


class _P(torch.nn.Module):
    def p(self, name, *shape, init=0.05):
        setattr(self, name, torch.nn.Parameter(torch.randn(*shape) * init))


class GN(_P):
    def __init__(self, c):
        super().__init__()
        self.gn = torch.nn.GroupNorm(32 if c % 32 == 0 else c // 2, c)

    def forward(self, x):
        return group_norm(x, self.gn.weight, self.gn.bias)


class GnReluConv(_P):
    def __init__(self, cin, cout, Fe, bias):
        super().__init__()
        self.norm = GN(cin)
        self.conv = _P()
        self.conv.p("weight", Fe * cin, cout)
        if bias:
            self.conv.p("bias", cout)

    def forward(self, x, table):
        y = lattice_conv(torch.relu(self.norm(x)), table, self.conv.weight)
        return y + self.conv.bias if hasattr(self.conv, "bias") else y


class GnRelu1x1(_P):
    def __init__(self, cin, cout, bias):
        super().__init__()
        self.norm = GN(cin)
        self.linear = torch.nn.Linear(cin, cout, bias=bias)

    def forward(self, x):
        return self.linear(torch.relu(self.norm(x)))


class ResnetBlock(_P):
    def __init__(self, c, Fe, biases):
        super().__init__()
        self.conv1 = GnReluConv(c, c, Fe, biases[0])
        self.conv2 = GnReluConv(c, c, Fe, biases[1])

    def forward(self, x, table):
        return self.conv2(self.conv1(x, table), table) + x


class BottleneckBlock(_P):
    def __init__(self, c, Fe, biases):
        super().__init__()
        self.contract = GnRelu1x1(c, c // 4, biases[0])
        self.conv = GnReluConv(c // 4, c // 4, Fe, biases[1])
        self.expand = GnRelu1x1(c // 4, c, biases[2])

    def forward(self, x, table):
        return self.expand(self.conv(self.contract(x), table)) + x


class CpuLNN(torch.nn.Module):
    def __init__(self, nr_classes, mp, pos_dim=3, val_dim=1):
        super().__init__()
        Fe = 2 * (pos_dim + 1) + 1
        self.mp, self.nr_classes, self.pos_dim = mp, nr_classes, pos_dim
        pn = list(mp.pointnet_channels_per_layer())
        start = mp.pointnet_start_nr_channels()
        self.point_net = _P()
        self.point_net.layers = torch.nn.ModuleList()
        cin = pos_dim + val_dim
        for cout in pn:
            layer = _P()
            layer.p("weight_g", cout, 1, init=1.0)
            layer.p("weight_v", cout, cin)
            layer.p("bias", cout)
            self.point_net.layers.append(layer)
            cin = cout
        self.point_net.last_conv = _P()
        self.point_net.last_conv.p("bias", start)
        self.point_net.last_conv.p("weight_g", 1, start, init=1.0)
        self.point_net.last_conv.p("weight_v", Fe * pn[-1] * 2, start)

        nd = mp.nr_downsamples()
        self.resnet_blocks_per_down_lvl_list = torch.nn.ModuleList()
        self.coarsens_list = torch.nn.ModuleList()
        skips, ch = [], start
        for lvl in range(nd):
            blocks = torch.nn.ModuleList()
            for _ in range(mp.nr_blocks_down_stage()[lvl]):
                blocks.append(ResnetBlock(ch, Fe, [False, False]) if lvl < mp.nr_levels_down_with_normal_resnet()
                              else BottleneckBlock(ch, Fe, [False] * 3))
            self.resnet_blocks_per_down_lvl_list.append(blocks)
            skips.append(ch)
            co = _P()
            co.coarse = _P()
            co.coarse.p("weight", Fe * ch, int(ch * 2 * mp.compression_factor()))
            self.coarsens_list.append(co)
            ch = int(ch * 2 * mp.compression_factor())
        self.resnet_blocks_bottleneck = torch.nn.ModuleList([BottleneckBlock(ch, Fe, [False] * 3) for _ in range(mp.nr_blocks_bottleneck())])
        self.finefy_list = torch.nn.ModuleList()
        self.resnet_blocks_per_up_lvl_list = torch.nn.ModuleList()
        for lvl in range(nd):
            skip = skips.pop()
            fi = _P()
            fi.norm = GN(ch)
            fi.fine = _P()
            fi.fine.p("weight", Fe * ch, ch // 2)
            self.finefy_list.append(fi)
            ch = skip + ch // 2
            blocks = torch.nn.ModuleList()
            for j in range(mp.nr_blocks_up_stage()[lvl]):
                last = j == mp.nr_blocks_up_stage()[lvl] - 1 and lvl == nd - 1
                blocks.append(ResnetBlock(ch, Fe, [False, last]) if lvl >= nd - mp.nr_levels_up_with_normal_resnet()
                              else BottleneckBlock(ch, Fe, [False, False, last]))
            self.resnet_blocks_per_up_lvl_list.append(blocks)
        s = _P()
        s.stepdown = torch.nn.ModuleList([GnRelu1x1(ch, ch, False), GnRelu1x1(ch, ch // 2, False)])
        s.bottleneck = GnRelu1x1(ch // 2, 8, False)
        s.linear_deltaW = torch.nn.Linear(9, 1)
        s.p("gamma", 9, init=1.0)
        s.p("beta", 9)
        s.linear_clasify = torch.nn.Linear(ch, nr_classes)
        self.slice_fast_cuda = s

    # ---------------------------------------------------------------------------------------------
    def structure(self, pos_np, sigmas, level1_keys=None):
        """Lattice levels for one cloud.  `level1_keys` fixes the level-1 vertex numbering (the reference's
        model treats vertex 0 as an 'invalid' row, lattice_modules.py:72-94,712, so numbering matters there)."""
        base = lo.build_lattice(pos_np, sigmas)
        idx, w = base["indices"], base["weights"]
        keys = base["keys"]
        if level1_keys is not None:
            look = {tuple(k): i for i, k in enumerate(np.asarray(level1_keys))}
            remap = np.array([look[tuple(k)] for k in keys], np.int64)
            idx = remap[idx]
            keys = np.asarray(level1_keys)
        levels = [Level(keys)]
        for l in range(1, self.mp.nr_downsamples() + 1):
            levels.append(Level(lo.build_lattice(pos_np, np.asarray(sigmas, np.float32) * np.float32(2 ** l))["keys"]))
        return levels, torch.from_numpy(idx.astype(np.int64)), torch.from_numpy(w), torch.from_numpy(base["scaled"])

    def forward(self, pos_np, values, sigmas, level1_keys=None):
        levels, idx, w, scaled = self.structure(pos_np, sigmas, level1_keys)
        n, d = scaled.shape
        spv = d + 1
        nv = levels[0].nv
        # ---- distribute + local mean subtraction (lattice_modules.py:52-96), no grad
        with torch.no_grad():
            dist_pos = scaled.repeat_interleave(spv, 0)
            sums = torch.zeros(nv, d).index_add_(0, idx, dist_pos)
            cnt = torch.zeros(nv).index_add_(0, idx, torch.ones(n * spv))
            mean = sums / cnt.clamp(min=1).unsqueeze(1)
            mean[0] = 0
            dist = torch.cat([dist_pos - mean[idx], values.repeat_interleave(spv, 0)], 1)
            dist = dist.masked_fill((idx == 0).unsqueeze(1), 0.0)
        # ---- PointNet (lattice_modules.py:661-733)
        x = dist
        for layer in self.point_net.layers:
            x = F_.leaky_relu(F_.linear(x, wn(layer.weight_v, layer.weight_g), layer.bias), 0.2)
        c = x.shape[1]
        big = torch.full((nv, c), -torch.inf).scatter_reduce(0, idx.unsqueeze(1).expand(-1, c), x, "amax", include_self=True)
        # argmax = smallest row attaining the max (ties are measure-zero for the real-valued features)
        hit = x == big[idx]
        rows = torch.where(hit, torch.arange(n * spv).unsqueeze(1).expand(-1, c), torch.full_like(hit, n * spv, dtype=torch.int64))
        arg = torch.full((nv, c), n * spv, dtype=torch.int64).scatter_reduce(0, idx.unsqueeze(1).expand(-1, c), rows, "amin", include_self=True)
        reduced = x.gather(0, arg.clamp(max=n * spv - 1))
        bary = w[arg.clamp(max=n * spv - 1)]
        reduced = torch.cat([reduced, bary], 1)
        reduced = reduced.masked_fill((cnt < 4).unsqueeze(1), 0.0)
        keep = torch.ones(nv, 1)
        keep[0] = 0
        reduced = reduced * keep
        lc = self.point_net.last_conv
        t_same = [lvl.table(lvl, 0) for lvl in levels]
        lv = F_.leaky_relu(lattice_conv(reduced, t_same[0], wn(lc.weight_v, lc.weight_g)) + lc.bias, 0.2)
        # ---- encoder
        skips = []
        nd = self.mp.nr_downsamples()
        for l in range(nd):
            for block in self.resnet_blocks_per_down_lvl_list[l]:
                lv = block(lv, t_same[l])
            skips.append(lv)
            lv = F_.leaky_relu(lattice_conv(lv, levels[l + 1].table(levels[l], 1), self.coarsens_list[l].coarse.weight), 0.2)
        for block in self.resnet_blocks_bottleneck:
            lv = block(lv, t_same[nd])
        # ---- decoder
        for i in range(nd):
            l = nd - 1 - i
            fi = self.finefy_list[i]
            lv = lattice_conv(torch.relu(fi.norm(lv)), levels[l].table(levels[l + 1], -1), fi.fine.weight)
            lv = torch.cat([lv, skips.pop()], 1)
            for block in self.resnet_blocks_per_up_lvl_list[i]:
                lv = block(lv, t_same[l])
        # ---- DeformSlice head (lattice_modules.py:465-567)
        s = self.slice_fast_cuda
        b = s.bottleneck(s.stepdown[1](s.stepdown[0](lv)))
        wi = w.reshape(n, spv)
        ii = idx.reshape(n, spv)
        gathered = torch.cat([b[ii] * wi.unsqueeze(2), wi.unsqueeze(2)], 2)           # [n, spv, 9]
        gathered = gathered - (s.gamma * gathered.max(1)[0].unsqueeze(1) + s.beta)
        dw = s.linear_deltaW(gathered).reshape(n, spv)
        sliced = (lv[ii] * (wi + dw).unsqueeze(2)).sum(1)
        logits = s.linear_clasify(sliced)
        return F_.log_softmax(logits, 1), logits


# --------------------------------------------------------------------------------------------------
def time_training_step(gpu_model, cloud_fn, nr_classes, sigma, budget_s=20.0):
    """cpu_baseline of bench.py: fwd + bwd + AdamW of the same architecture on the host cores, over a
    bounded sample of the same synthetic clouds."""
    from lattice_net_b200.losses import segmentation_loss
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    mp = gpu_model.model_params
    model = CpuLNN(nr_classes, mp)
    try:
        model.load_state_dict({k: v.detach().cpu() for k, v in gpu_model.state_dict().items()}, strict=True)
    except Exception:
        pass    # timing does not depend on the parameter values
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=3e-4, amsgrad=True)
    done, t_total = 0, 0.0
    t_start = time.time()
    i = 0
    while True:
        pos, labels = cloud_fn(50_000 + i)
        t0 = time.time()
        logsm, _ = model(pos, torch.zeros(pos.shape[0], 1), [sigma] * 3)
        loss = segmentation_loss(logsm, torch.from_numpy(labels))
        opt.zero_grad()
        loss.backward()
        opt.step()
        dt = time.time() - t0
        if i > 0:            # first scan warms the allocator / thread pool
            done += 1
            t_total += dt
        i += 1
        if (time.time() - t_start > budget_s and done >= 2) or done >= 80:     # ~15-20 s of host work either way
            break
    return {"value": done / t_total, "unit": "scans/s", "cores": cores, "kind": "port",
            "sample": f"{done} scans of the same workload (2048-pt clouds, fwd+bwd+AdamW), torch-CPU gather/index_add/mm port incl. lattice construction on 1 core (C oracle)",
            "ms_per_scan": 1e3 * t_total / done}
