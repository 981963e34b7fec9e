"""CPU-only tests of the host logic: cfg reader, parameter structs, C-ABI library symbols."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CFG = '''
core: { loguru_verbosity: 3  hidpi: false }
train: { dataset_name: "shapenet" // comment
    lr: 0.001
    weight_decay: 3e-4
    save_checkpoint: false
    checkpoint_path: ""
}
model: {
    positions_mode: "xyz"
    values_mode: "none"
    pointnet_layers: [16,32,64]
    pointnet_start_nr_channels: 32
    nr_downsamples: 3
    nr_blocks_down_stage: [6,6,8]
    nr_blocks_bottleneck: 8
    nr_blocks_up_stage: [2,2,2]
    nr_levels_down_with_normal_resnet: 3
    nr_levels_up_with_normal_resnet: 3
    compression_factor: 1.0
    dropout_last_layer: 0.0
}
lattice_gpu: {
    hash_table_capacity: 60000 //good for shapenet
    nr_sigmas: 2
    // sigma_0: "0.06 3"
    sigma_0: "0.05 3"
    sigma_1: "0.1 2"
}
'''


def test_cfg_parser_and_params():
    from lattice_net_b200 import params
    cfg = params.parse_cfg_text(CFG)
    assert params.lattice_settings(cfg) == (60000, [(0.05, 3), (0.1, 2)])
    mp = params.ModelParams.create(cfg)
    assert mp.pointnet_channels_per_layer() == [16, 32, 64]      # old key spelling accepted
    assert mp.nr_blocks_down_stage() == [6, 6, 8] and mp.nr_blocks_bottleneck() == 8
    tp = params.TrainParams.create(cfg)
    assert tp.lr() == 0.001 and tp.weight_decay() == 3e-4 and tp.dataset_name() == "shapenet" and tp.save_checkpoint() is False
    assert params.EvalParams.create(cfg).do_write_predictions() is False
    with pytest.raises(ValueError):
        params.parse_cfg_text("a: { b: 1 ")
    with pytest.raises(FileNotFoundError):
        params.parse_cfg("does_not_exist.cfg")


def test_default_model_params_are_the_shapenet_architecture():
    from lattice_net_b200 import ModelParams
    mp = ModelParams()
    assert mp.pointnet_start_nr_channels() == 32 and mp.nr_downsamples() == 3
    assert mp.nr_blocks_down_stage() == [3, 3, 3] and mp.nr_blocks_up_stage() == [2, 2, 2]


def test_lattice_handle_host_logic():
    from lattice_net_b200 import Lattice
    lat = Lattice(1000, [(0.05, 3)], name="l")
    assert Lattice.get_expected_filter_extent(1) == 9 and lat.name() == "l" and lat.capacity() == 1000
    assert lat.m_sigmas == [0.05, 0.05, 0.05]
    lat.increase_sigmas(0.01)
    assert abs(lat.m_sigmas[0] - 0.06) < 1e-9
    lat.set_sigma(0.1)
    assert lat.m_sigmas == [0.1] * 3
    with pytest.raises(RuntimeError):
        lat.nr_lattice_vertices()            # nothing splatted yet
    with pytest.raises(RuntimeError):
        Lattice.get_expected_filter_extent(2)
    lat5 = Lattice(10, [(0.1, 3), (0.2, 2)])
    assert Lattice.get_expected_filter_extent(1) == 13 and len(lat5.m_sigmas) == 5
    clone = lat5.clone_lattice()
    assert clone.m_sigmas == lat5.m_sigmas and clone.lvl() == 1 and clone.hash_table().structure is lat5.hash_table().structure
    Lattice(1000, [(0.05, 3)])               # restore the static expected pos_dim for other tests


def test_c_abi_library_exports_every_declared_symbol():
    """include/lattice_b200.h is the contract: every declared ln_* function must be exported."""
    from lattice_net_b200 import _cabi, build
    if not os.path.isfile(_cabi.LIB_PATH):
        build.build()
    with open(os.path.join(ROOT, "include", "lattice_b200.h")) as f:
        header = f.read()
    declared = sorted(set(re.findall(r"\b(ln_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 25
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
    assert sorted(_cabi.EXPORTED_SYMBOLS) == declared, "ctypes signature table and header disagree"
    _cabi.load()
    assert _cabi.version().endswith("sm_100a")


def test_workspace_size_queries_are_host_only_and_consistent():
    """ln_conv_workspace_bytes / ln_group_norm_workspace_bytes are pure host functions (callable without a GPU); the
    Python-side copies of their rules (lattice.py:_slab_floats / tensor_core_reading_ok, lattice_modules.py:_gn_workspace)
    must agree."""
    from lattice_net_b200 import _cabi, lattice as lattice_mod
    lib = _cabi.load()
    F = 9
    saved = lattice_mod.CONV_PRECISION
    try:
        lattice_mod.CONV_PRECISION = 1
        for c_in, c_out in [(32, 32), (64, 128), (128, 96), (256, 256), (64, 512), (512, 384), (96, 7), (3, 32), (32, 2048)]:
            want = int(lib.ln_conv_workspace_bytes(F, c_in, c_out, 1))
            mine = 4 * lattice_mod._slab_floats(F, c_in, c_out) if lattice_mod.tensor_core_reading_ok(F, c_in, c_out) else 0
            assert mine == want, (c_in, c_out)
        assert int(lib.ln_conv_workspace_bytes(1, 128, 64, 1)) == 2 * 128 * 64 * 4     # filter extent 1: the 1x1 layers
        # layers wider than one 256-column tile are chunked, not sent to the fp32 kernel
        assert int(lib.ln_conv_workspace_bytes(F, 64, 512, 1)) == 2 * F * 64 * 512 * 4
        assert int(lib.ln_conv_workspace_bytes(F, 3, 32, 1)) == 0          # c_in % 32 != 0: fp32 kernel, no workspace
        assert int(lib.ln_conv_workspace_bytes(F, 64, 64, 0)) == 0         # precision 0
    finally:
        lattice_mod.CONV_PRECISION = saved
    gn = lib.ln_group_norm_workspace_bytes
    assert int(gn(1000, 128, 32)) == 0                       # ShapeNet-sized level: one CTA per group, no scratch
    assert int(gn(673, 512, 32)) > 0                         # 16 channels per group: row-tiled even when small
    big = int(gn(34809, 64, 32))
    rows_per_cta = (256 // (64 // 4)) * 8
    ctas = -(-34809 // rows_per_cta)
    assert big == (ctas * 2 * 64 + 148 * 2 * 64 + 2 * 32) * 4  # per-CTA partials + pre-reduced partials + per-group (ds, db)
    assert int(gn(34809, 66, 33)) == 0                       # C % 4 != 0: generic kernel


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from lattice_net_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_cabi.LatticeBackendError, match="no CPU or PyTorch fallback"):
        _cabi.load()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "lattice_net_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            with open(os.path.join(pkg, fn)) as f:
                src = f.read()
            assert "oracle" not in src.replace("# oracle", ""), f"{fn} mentions the oracle"


def test_batched_lovasz_equals_per_class_loop():
    import torch
    from lattice_net_b200.losses import lovasz_softmax, lovasz_softmax_loop, segmentation_loss
    torch.manual_seed(0)
    logits = torch.randn(500, 7, requires_grad=True)
    labels = torch.randint(0, 5, (500,))            # classes 5 and 6 absent
    p = torch.softmax(logits, 1)
    a, b = lovasz_softmax(p, labels), lovasz_softmax_loop(p, labels)
    assert torch.allclose(a, b, atol=1e-6)
    ga, = torch.autograd.grad(a, logits, retain_graph=True)
    gb, = torch.autograd.grad(b, logits)
    assert torch.allclose(ga, gb, atol=1e-6)
    assert torch.isfinite(segmentation_loss(torch.log_softmax(logits, 1), labels))
    # ignore_index leaves that CLASS out of the mean but keeps its points as negatives of the other classes
    # (the reference's class loop, lovasz_loss.py:44-45), value and gradient
    p = torch.softmax(logits, 1)
    a, b = lovasz_softmax(p, labels, ignore_index=2), lovasz_softmax_loop(p, labels, ignore_index=2)
    assert torch.allclose(a, b, atol=1e-6)
    ga, = torch.autograd.grad(a, logits, retain_graph=True)
    gb, = torch.autograd.grad(b, logits)
    assert torch.allclose(ga, gb, atol=1e-6)


# Names the reference binds to Python (live `.def` / `.def_static` / `.def_readonly` lines of
# /root/reference/src/PyBridge.cxx:27-154, extracted once; the test does not read the reference tree).
REFERENCE_PYTHON_SURFACE = {
    "HashTable": [
        "m_keys_tensor",
        "m_nr_filled_tensor"
    ],
    "Lattice": [
        "begin_splat",
        "capacity",
        "clone_lattice",
        "convolve_im2row_standalone",
        "create",
        "create_coarse_verts",
        "create_coarse_verts_naive",
        "distribute",
        "expand",
        "gather_backwards_standalone_with_precomputation",
        "gather_standalone_no_precomputation",
        "gather_standalone_with_precomputation",
        "get_expected_filter_extent",
        "get_filter_extent",
        "hash_table",
        "im2row",
        "im2rowindices",
        "increase_sigmas",
        "just_create_verts",
        "name",
        "nr_lattice_vertices",
        "pos_dim",
        "positions",
        "row2im",
        "set_positions",
        "set_sigma",
        "set_values",
        "sigmas_tensor",
        "slice_backwards_standalone_with_precomputation",
        "slice_backwards_standalone_with_precomputation_no_homogeneous",
        "slice_classify_backwards_with_precomputation",
        "slice_classify_no_precomputation",
        "slice_classify_with_precomputation",
        "slice_standalone_no_precomputation",
        "slice_standalone_with_precomputation",
        "splat_standalone",
        "val_dim",
        "values"
    ],
    "TrainParams": [
        "checkpoint_path",
        "create",
        "dataset_name",
        "lr",
        "save_checkpoint",
        "weight_decay",
        "with_tensorboard",
        "with_viewer",
        "with_visdom"
    ],
    "EvalParams": [
        "checkpoint_path",
        "create",
        "dataset_name",
        "do_write_predictions",
        "output_predictions_path",
        "with_viewer"
    ],
    "ModelParams": [
        "compression_factor",
        "create",
        "dropout_last_layer",
        "nr_blocks_bottleneck",
        "nr_blocks_down_stage",
        "nr_blocks_up_stage",
        "nr_downsamples",
        "nr_levels_down_with_normal_resnet",
        "nr_levels_up_with_normal_resnet",
        "pointnet_channels_per_layer",
        "pointnet_start_nr_channels",
        "positions_mode",
        "values_mode"
    ]
}


def test_python_surface_of_the_reference_is_kept():
    """Drop-in boundary: every class / method / attribute name the reference's pybind module exposes exists here."""
    import lattice_net_b200 as pkg
    lattice = pkg.Lattice(60000, [(0.05, 3)])
    for cls_name, names in REFERENCE_PYTHON_SURFACE.items():
        cls = getattr(pkg, cls_name, None)
        assert cls is not None, f"class {cls_name} is missing"
        obj = lattice if cls_name == "Lattice" else cls
        missing = [n for n in names if not hasattr(obj, n)]
        assert not missing, f"{cls_name} lacks {missing}"


def test_input_path_file_formats(tmp_path):
    """the on-disk formats either side of the path (ScanNet .ply, SemanticKITTI .bin / .label, prediction files)
    and prepare_cloud's positions / values modes (reference models.py:18-66) -- host code, no GPU."""
    import types
    import numpy as np
    import torch
    from lattice_net_b200 import data

    rng = np.random.RandomState(3)
    n = 257
    pos = rng.randn(n, 3).astype(np.float32)
    col = rng.randint(0, 256, (n, 3)).astype(np.float32) / 255.0
    lab = rng.randint(0, 21, n)
    for binary in (True, False):
        p = str(tmp_path / ("cloud_%d.ply" % binary))
        data.write_ply_cloud(p, pos, col, lab, binary=binary)
        c = data.read_ply_cloud(p)
        assert c.V.dtype == np.float32 and np.array_equal(c.V, pos)          # float32 round-trips exactly (repr in ascii)
        assert np.allclose(c.C, col, atol=0.5 / 255.0 + 1e-7) and c.C.max() <= 1.0
        assert np.array_equal(c.L_gt.reshape(-1), lab) and c.L_gt.shape == (n, 1)
        assert c.name == "cloud_%d" % binary
    # ScanNet layout: geometry + colour (+ alpha, + faces) in one file, labels in a second file with the same vertices
    mesh = str(tmp_path / "scene0000_00_vh_clean_2.ply")
    rec = np.zeros(n, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("red", "u1"), ("green", "u1"), ("blue", "u1"), ("alpha", "u1")])
    rec["x"], rec["y"], rec["z"] = pos.T
    rec["red"], rec["green"], rec["blue"] = np.rint(col * 255).astype(np.uint8).T
    with open(mesh, "wb") as f:
        f.write(("ply\nformat binary_little_endian 1.0\ncomment VCGLIB generated\nelement vertex %d\nproperty float x\nproperty float y\n"
                 "property float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nproperty uchar alpha\nelement face 1\n"
                 "property list uchar int vertex_indices\nend_header\n" % n).encode())
        f.write(rec.tobytes())
        f.write(bytes([3]) + np.array([0, 1, 2], "<i4").tobytes())
    labels = str(tmp_path / "scene0000_00_vh_clean_2.labels.ply")
    data.write_ply_cloud(labels, pos, col, lab)
    c = data.read_ply_cloud(mesh, labels)
    assert np.array_equal(c.V, pos) and np.array_equal(c.L_gt.reshape(-1), lab) and np.allclose(c.C, col, atol=1e-6)
    assert np.array_equal(data.read_ply_cloud(mesh).L_gt, np.zeros((n, 1), np.int32))
    with pytest.raises(ValueError):
        data.write_ply_cloud(str(tmp_path / "short.ply"), pos[:5], None, lab[:5])
        data.read_ply_cloud(mesh, str(tmp_path / "short.ply"))
    with open(str(tmp_path / "trunc.ply"), "wb") as f:
        f.write(open(mesh, "rb").read()[:400])
    with pytest.raises(ValueError):
        data.read_ply_cloud(str(tmp_path / "trunc.ply"))
    with pytest.raises(ValueError):
        data.read_ply_cloud(__file__)

    # SemanticKITTI: .bin float32 x y z remission, .label uint32 with the instance id in the upper 16 bits
    scan = rng.randn(100, 4).astype(np.float32)
    sem = rng.randint(0, 20, 100).astype(np.uint32)
    (scan).tofile(str(tmp_path / "000000.bin"))
    (sem | (rng.randint(0, 1000, 100).astype(np.uint32) << 16)).tofile(str(tmp_path / "000000.label"))
    k = data.read_semantic_kitti_scan(str(tmp_path / "000000.bin"), str(tmp_path / "000000.label"))
    assert np.array_equal(k.V, scan[:, :3]) and np.array_equal(k.I, scan[:, 3:4]) and np.array_equal(k.L_gt.reshape(-1), sem.astype(np.int32))
    assert np.array_equal(data.read_semantic_kitti_scan(str(tmp_path / "000000.bin")).L_gt, np.zeros((100, 1), np.int32))

    # prediction writers (ln_eval.py:160-191)
    logits = torch.from_numpy(rng.randn(100, 20).astype(np.float32))
    out = data.write_label_file(logits, str(tmp_path / "pred" / "000000.label"))
    back = np.fromfile(str(tmp_path / "pred" / "000000.label"), dtype=np.uint32)
    assert back.dtype == np.uint32 and np.array_equal(back, logits.argmax(1).numpy()) and np.array_equal(out, back)
    ids = np.arange(20) * 2 + 1
    data.write_scannet_evaluation_file(logits, str(tmp_path / "eval" / "scene0000_00.txt"), ids)
    txt = np.loadtxt(str(tmp_path / "eval" / "scene0000_00.txt"), dtype=np.int64)
    assert np.array_equal(txt, ids[logits.argmax(1).numpy()])

    # prepare_cloud: every positions / values mode of reference models.py:18-66
    cloud = types.SimpleNamespace(V=pos, C=col.astype(np.float32), I=rng.rand(n, 1).astype(np.float32), L_gt=lab.reshape(-1, 1).astype(np.int32))

    def mp(pm, vm):
        return types.SimpleNamespace(positions_mode=lambda: pm, values_mode=lambda: vm)

    T = torch.from_numpy
    want_pos = {"xyz": T(pos), "xyz+rgb": torch.cat((T(pos), T(cloud.C)), 1), "xyz+intensity": torch.cat((T(pos), T(cloud.I)), 1)}
    want_val = {"none": torch.zeros(n, 1), "intensity": T(cloud.I), "rgb": T(cloud.C), "rgb+height": torch.cat((T(cloud.C), T(pos[:, 1:2].copy())), 1),
                "rgb+xyz": torch.cat((T(cloud.C), T(pos)), 1), "height": T(pos[:, 1:2].copy()), "xyz": T(pos)}
    for pm, wp in want_pos.items():
        for vm, wv in want_val.items():
            p_, v_, t_ = data.prepare_cloud(cloud, mp(pm, vm), device="cpu")
            assert torch.equal(p_, wp) and torch.equal(v_, wv) and p_.is_contiguous() and v_.is_contiguous()
            assert t_.dtype == torch.int64 and torch.equal(t_, T(lab.astype(np.int64)))
    as_dict = {"V": cloud.V, "C": cloud.C, "I": cloud.I, "L_gt": cloud.L_gt}
    assert torch.equal(data.prepare_cloud(as_dict, mp("xyz", "rgb"), device="cpu")[1], T(cloud.C))
    with pytest.raises(SystemExit):
        data.prepare_cloud(cloud, mp("uvw", "none"), device="cpu")
    with pytest.raises(SystemExit):
        data.prepare_cloud(cloud, mp("xyz", "normals"), device="cpu")
