"""Input path of the training / evaluation scripts.

  * `prepare_cloud(cloud, model_params)` -- /root/reference/latticenet_py/lattice/models.py:18-66: positions / values /
    target tensors of a cloud according to `positions_mode` and `values_mode` of the model config.  `cloud` is anything
    with the reference's mesh attributes as numpy arrays: V [N x 3] positions, C [N x 3] colours, I [N x 1] intensity,
    L_gt [N x 1] labels (EasyPBR's Mesh has exactly these; a types.SimpleNamespace or a dict works as well).
  * `SyntheticCloud`, `read_semantic_kitti_scan`, `write_label_file` -- the SemanticKITTI on-disk formats the reference's
    loader / eval script handle (`.bin` float32 x,y,z,intensity; `.label` uint32, lower 16 bits = class;
    /root/reference/latticenet_py/ln_eval.py:168-193 writes predictions as uint32 `.label` files).
  * `read_ply_cloud`, `write_ply_cloud` -- the ScanNet on-disk format (`*_vh_clean_2.ply` / `*_vh_clean_2.labels.ply`: PLY vertex
    elements x, y, z float, red, green, blue, alpha uchar, optional label ushort; ascii or binary), and the `_pred.ply` /
    `_gt.ply` files ln_eval.py:143-147 names; `write_scannet_evaluation_file`: the benchmark's one-label-id-per-line text file
    (ln_eval.py:160-163 hands this to the loader's `write_for_evaluating_on_scannet_server`).
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


_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
              "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}


def _read_ply_vertices(path):
    """Vertex element of a PLY file as a numpy structured array (other elements, e.g. faces, are not read)."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt = None
        elements = []          # (name, count, [(prop name, dtype) | None for list properties])
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: PLY header without end_header")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] in ("comment", "obj_info"):
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                elements.append((tok[1], int(tok[2]), []))
            elif tok[0] == "property":
                if tok[1] == "list":
                    elements[-1][2].append(None)
                else:
                    if tok[1] not in _PLY_TYPES:
                        raise ValueError(f"{path}: PLY property type {tok[1]} unknown")
                    elements[-1][2].append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
            raise ValueError(f"{path}: PLY format {fmt} unknown")
        if not elements or elements[0][0] != "vertex":
            raise ValueError(f"{path}: the first PLY element must be `vertex`")
        _, count, props = elements[0]
        if any(p is None for p in props):
            raise ValueError(f"{path}: list properties on vertices are not supported")
        order = "<" if fmt != "binary_big_endian" else ">"
        dtype = np.dtype([(n, order + t) for n, t in props])
        if fmt == "ascii":
            rows = np.loadtxt(f, dtype=np.float64, max_rows=count, ndmin=2) if count else np.zeros((0, len(props)))
            if rows.shape != (count, len(props)):
                raise ValueError(f"{path}: expected {count} vertex rows of {len(props)} values")
            out = np.zeros(count, dtype=dtype)
            for i, (n, _) in enumerate(props):
                out[n] = rows[:, i]
            return out
        raw = f.read(count * dtype.itemsize)
        if len(raw) != count * dtype.itemsize:
            raise ValueError(f"{path}: truncated PLY vertex data")
        return np.frombuffer(raw, dtype=dtype, count=count)


def read_ply_cloud(path, labels_path=None):
    """A cloud (V, C in [0,1], I zeros, L_gt) from a PLY file; per-vertex labels come from the `label` property of `labels_path`
    (ScanNet keeps them in a second file with the same vertices) or of `path` itself, else zeros."""
    v = _read_ply_vertices(path)
    names = v.dtype.names
    n = len(v)
    c = types.SimpleNamespace(V=np.stack([v["x"], v["y"], v["z"]], 1).astype(np.float32) if n else np.zeros((0, 3), np.float32))
    if all(k in names for k in ("red", "green", "blue")):
        rgb = np.stack([v["red"], v["green"], v["blue"]], 1)
        c.C = (rgb.astype(np.float32) / 255.0) if rgb.dtype.kind in "ui" else rgb.astype(np.float32)
    else:
        c.C = np.zeros((n, 3), np.float32)
    c.I = np.zeros((n, 1), np.float32)
    lab = None
    if labels_path is not None:
        lv = _read_ply_vertices(labels_path)
        if len(lv) != n:
            raise ValueError(f"{labels_path}: {len(lv)} vertices, the cloud has {n}")
        if "label" not in lv.dtype.names:
            raise ValueError(f"{labels_path}: no `label` vertex property")
        lab = lv["label"]
    elif "label" in names:
        lab = v["label"]
    c.L_gt = (lab.astype(np.int32) if lab is not None else np.zeros(n, np.int32)).reshape(-1, 1)
    c.m_disk_path = path
    c.name = os.path.splitext(os.path.basename(path))[0]
    return c


def write_ply_cloud(path, positions, colours=None, labels=None, binary=True):
    """Positions [N x 3] (+ colours in [0,1] [N x 3], + labels [N]) as a PLY vertex cloud -- the `_pred.ply` / `_gt.ply` of ln_eval.py."""
    pos = np.asarray(positions, dtype=np.float32).reshape(-1, 3)
    fields = [("x", "<f4"), ("y", "<f4"), ("z", "<f4")]
    header = ["ply", "format %s 1.0" % ("binary_little_endian" if binary else "ascii"), "element vertex %d" % len(pos),
              "property float x", "property float y", "property float z"]
    if colours is not None:
        fields += [("red", "u1"), ("green", "u1"), ("blue", "u1")]
        header += ["property uchar red", "property uchar green", "property uchar blue"]
    if labels is not None:
        fields += [("label", "<u2")]
        header += ["property ushort label"]
    header.append("end_header")
    rec = np.zeros(len(pos), dtype=np.dtype(fields))
    rec["x"], rec["y"], rec["z"] = pos[:, 0], pos[:, 1], pos[:, 2]
    if colours is not None:
        col = np.clip(np.rint(np.asarray(colours, dtype=np.float32).reshape(-1, 3) * 255.0), 0, 255).astype(np.uint8)
        rec["red"], rec["green"], rec["blue"] = col[:, 0], col[:, 1], col[:, 2]
    if labels is not None:
        rec["label"] = np.asarray(labels).reshape(-1).astype(np.uint16)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        if binary:
            f.write(rec.tobytes())
        else:
            for r in rec:
                f.write((" ".join(repr(x.item()) if isinstance(x, np.floating) else str(x.item()) for x in r) + "\n").encode("ascii"))


def write_scannet_evaluation_file(pred_logsoftmax, path, class_to_benchmark_id=None):
    """One label id per line, in vertex order (the ScanNet benchmark's submission format).  `class_to_benchmark_id` maps the
    network's class index to the benchmark's id (e.g. the NYU40 ids of the 20 evaluated classes); identity when None."""
    l_pred = pred_logsoftmax.detach().argmax(dim=1).cpu().numpy().reshape(-1).astype(np.int64)
    if class_to_benchmark_id is not None:
        l_pred = np.asarray(class_to_benchmark_id, dtype=np.int64)[l_pred]
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    np.savetxt(path, l_pred, fmt="%d")
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
