"""Input path of the training / evaluation scripts.

  * `prepare_cloud(cloud, model_params)` -- /root/reference/latticenet_py/lattice/models.py:18-66: positions / values /
    target tensors of a cloud according to `positions_mode` and `values_mode` of the model config.  `cloud` is anything
    with the reference's mesh attributes as numpy arrays: V [N x 3] positions, C [N x 3] colours, I [N x 1] intensity,
    L_gt [N x 1] labels (EasyPBR's Mesh has exactly these; a types.SimpleNamespace or a dict works as well).
  * `SyntheticCloud`, `read_semantic_kitti_scan`, `write_label_file` -- the SemanticKITTI on-disk formats the reference's
    loader / eval script handle (`.bin` float32 x,y,z,intensity; `.label` uint32, lower 16 bits = class;
    /root/reference/latticenet_py/ln_eval.py:168-193 writes predictions as uint32 `.label` files).
  * `PinnedCloudFeeder` -- double-buffered pinned-memory H2D staging for the graphed step: while the GPU replays the step
    graph of cloud i, cloud i+1 is copied into the other set of device buffers on a copy stream, so the end-to-end rate
    does not pay the host-to-device copy.
"""
import os
import sys
import types

import numpy as np
import torch


def _field(cloud, name):
    v = cloud[name] if isinstance(cloud, dict) else getattr(cloud, name)
    return np.asarray(v)


def _f32(a, device):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device)


def prepare_cloud(cloud, model_params, device="cuda"):
    """-> (positions [N x pos_dim], values [N x val_dim], target int64 [N]) on `device`."""
    with torch.no_grad():
        pm = model_params.positions_mode()
        if pm == "xyz":
            positions = _f32(_field(cloud, "V"), device)
        elif pm == "xyz+rgb":
            positions = torch.cat((_f32(_field(cloud, "V"), device), _f32(_field(cloud, "C"), device)), 1)
        elif pm == "xyz+intensity":
            positions = torch.cat((_f32(_field(cloud, "V"), device), _f32(_field(cloud, "I"), device)), 1)
        else:
            sys.exit(f"positions mode of {pm} not implemented")
        vm = model_params.values_mode()
        if vm == "none":
            values = torch.zeros((positions.shape[0], 1), device=device)     # the lattice needs some value array
        elif vm == "intensity":
            values = _f32(_field(cloud, "I"), device)
        elif vm == "rgb":
            values = _f32(_field(cloud, "C"), device)
        elif vm == "rgb+height":
            values = torch.cat((_f32(_field(cloud, "C"), device), _f32(_field(cloud, "V")[:, 1:2], device)), 1)
        elif vm == "rgb+xyz":
            values = torch.cat((_f32(_field(cloud, "C"), device), _f32(_field(cloud, "V"), device)), 1)
        elif vm == "height":
            values = _f32(_field(cloud, "V")[:, 1:2], device)
        elif vm == "xyz":
            values = _f32(_field(cloud, "V"), device)
        else:
            sys.exit(f"values mode of {vm} not implemented")
        target = torch.from_numpy(np.ascontiguousarray(_field(cloud, "L_gt")).astype(np.int64).reshape(-1)).to(device)
    return positions.contiguous(), values.contiguous(), target


def SyntheticCloud(n, nr_classes, seed, with_colour=False, with_intensity=False):
    """A cloud object with the reference's mesh attributes (V, C, I, L_gt), points on the faces of a box."""
    rng = np.random.RandomState(seed)
    size = np.array([0.8, 0.3, 0.4])
    p = (rng.rand(n, 3) - 0.5) * size
    face = rng.randint(0, 3, n)
    p[np.arange(n), face] = 0.5 * size[face] * (rng.randint(0, 2, n) * 2 - 1)
    c = types.SimpleNamespace(V=p.astype(np.float32), L_gt=rng.randint(0, nr_classes, (n, 1)).astype(np.int32))
    c.C = rng.rand(n, 3).astype(np.float32) if with_colour else np.zeros((n, 3), np.float32)
    c.I = rng.rand(n, 1).astype(np.float32) if with_intensity else np.zeros((n, 1), np.float32)
    return c


def read_semantic_kitti_scan(bin_path, label_path=None):
    """SemanticKITTI velodyne scan: float32 [N x 4] (x, y, z, remission); labels uint32, lower 16 bits = semantic class."""
    scan = np.fromfile(bin_path, dtype=np.float32).reshape(-1, 4)
    c = types.SimpleNamespace(V=np.ascontiguousarray(scan[:, :3]), I=np.ascontiguousarray(scan[:, 3:4]), C=np.zeros((len(scan), 3), np.float32))
    if label_path is not None and os.path.isfile(label_path):
        lab = np.fromfile(label_path, dtype=np.uint32)
        c.L_gt = (lab & 0xFFFF).astype(np.int32).reshape(-1, 1)
    else:
        c.L_gt = np.zeros((len(scan), 1), np.int32)
    c.m_disk_path = bin_path
    return c


def write_label_file(pred_logsoftmax, path):
    """ln_eval.py:186-191: per-point argmax as a uint32 `.label` file."""
    l_pred = pred_logsoftmax.detach().argmax(dim=1).cpu().numpy().reshape(-1).astype(np.uint32)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    l_pred.tofile(path)
    return l_pred


class PinnedCloudFeeder:
    """Two sets of (positions, values, labels) device buffers fed from pinned host memory on a copy stream.

        feeder = PinnedCloudFeeder(nr_points, pos_dim, val_dim, device)
        feeder.stage(pos_np, vals_np, labels_np)            # cloud 0
        for next_cloud in clouds[1:] + [None]:
            slot, (pos, vals, labels) = feeder.current()    # the compute stream waits for the copy of this cloud
            if next_cloud is not None:
                feeder.stage(*next_cloud)                   # H2D of the next cloud runs under this step
            step(pos, vals, labels)
            feeder.release(slot)                            # from here on in stream order the slot may be overwritten
    """

    def __init__(self, nr_points, pos_dim, val_dim, device):
        self.device = device
        self.copy_stream = torch.cuda.Stream(device=device)
        self.slots = []
        for _ in range(2):
            host = (torch.empty((nr_points, pos_dim), dtype=torch.float32).pin_memory(), torch.empty((nr_points, val_dim), dtype=torch.float32).pin_memory(),
                    torch.empty((nr_points,), dtype=torch.int64).pin_memory())
            dev = tuple(torch.empty_like(h, device=device) for h in host)
            self.slots.append({"host": host, "dev": dev, "ready": torch.cuda.Event(), "free": torch.cuda.Event()})
        self.staged = -1     # slot holding the most recently staged cloud
        self.used = set()

    def stage(self, positions, values, labels):
        slot = (self.staged + 1) % 2
        s = self.slots[slot]
        if slot in self.used:
            s["free"].synchronize()                      # the step that read this slot's device buffers has finished with them
        for h, src in zip(s["host"], (positions, values, labels)):
            h.copy_(torch.as_tensor(src).reshape(h.shape))
        with torch.cuda.stream(self.copy_stream):
            for d, h in zip(s["dev"], s["host"]):
                d.copy_(h, non_blocking=True)
            s["ready"].record(self.copy_stream)
        self.staged = slot
        return slot

    def current(self):
        """(slot, device tensors) of the most recently staged cloud; the compute stream waits for their copy."""
        slot = self.staged
        s = self.slots[slot]
        torch.cuda.current_stream(self.device).wait_event(s["ready"])
        return slot, s["dev"]

    def release(self, slot):
        """Call after launching the work that consumes slot's tensors: marks (in stream order) when it may be overwritten."""
        self.slots[slot]["free"].record(torch.cuda.current_stream(self.device))
        self.used.add(slot)
