"""Reader for the reference's configuru `.cfg` files plus the three parameter structs the reference
binds to Python (`TrainParams`, `ModelParams`, `EvalParams`, /root/reference/src/PyBridge.cxx:116-154;
key names from src/TrainParams.cxx, src/ModelParams.cxx, src/EvalParams.cxx).

The cfg dialect is relaxed JSON: `key: value` pairs, `{}` blocks, `[]` lists, `//` comments, optional
commas, quoted strings, bare numbers / true / false.
"""
import os
import re

_TOKEN = re.compile(r'\s*(?:(//[^\n]*)|("(?:[^"\\]|\\.)*")|([{}\[\]:,])|([^\s{}\[\]:,"]+))')


def _tokenize(text):
    pos, out = 0, []
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            if text[pos:].strip() == "":
                break
            raise ValueError(f"cfg: cannot tokenize near {text[pos:pos + 30]!r}")
        pos = m.end()
        comment, string, punct, bare = m.groups()
        if comment is not None:
            continue
        if string is not None:
            out.append(("str", bytes(string[1:-1], "utf-8").decode("unicode_escape")))
        elif punct is not None:
            out.append(("punct", punct))
        else:
            out.append(("bare", bare))
    return out


def _scalar(tok):
    kind, val = tok
    if kind == "str":
        return val
    low = val.lower()
    if low == "true":
        return True
    if low == "false":
        return False
    try:
        return int(val)
    except ValueError:
        try:
            return float(val)
        except ValueError:
            return val


def _parse_value(toks, i):
    kind, val = toks[i]
    if kind == "punct" and val == "{":
        return _parse_block(toks, i + 1)
    if kind == "punct" and val == "[":
        items, i = [], i + 1
        while toks[i] != ("punct", "]"):
            if toks[i] == ("punct", ","):
                i += 1
                continue
            v, i = _parse_value(toks, i)
            items.append(v)
        return items, i + 1
    return _scalar(toks[i]), i + 1


def _parse_block(toks, i, top=False):
    out = {}
    while i < len(toks):
        if toks[i] == ("punct", "}"):
            if top:
                raise ValueError("cfg: unbalanced '}'")
            return out, i + 1
        if toks[i] == ("punct", ","):
            i += 1
            continue
        key = toks[i][1]
        if i + 1 >= len(toks) or toks[i + 1] != ("punct", ":"):
            raise ValueError(f"cfg: expected ':' after key {key!r}")
        out[key], i = _parse_value(toks, i + 2)
    if not top:
        raise ValueError("cfg: missing '}'")
    return out, i


def parse_cfg_text(text):
    cfg, _ = _parse_block(_tokenize(text), 0, top=True)
    return cfg


def parse_cfg(path, search_dirs=()):
    """Relative paths are tried as given, then under each of `search_dirs`, then under
    $LATTICENET_CONFIG_DIR (the reference resolves them against its compile-time PROJECT_SOURCE_DIR,
    /root/reference/src/Lattice.cu:109-115)."""
    cands = [path]
    if not os.path.isabs(path):
        for d in list(search_dirs) + [os.environ.get("LATTICENET_CONFIG_DIR", "")]:
            if d:
                cands += [os.path.join(d, path), os.path.join(d, "config", path)]
    for c in cands:
        if os.path.isfile(c):
            with open(c) as f:
                return parse_cfg_text(f.read())
    raise FileNotFoundError(f"config file {path!r} not found (tried {cands})")


def lattice_settings(cfg):
    """`lattice_gpu` block -> (capacity, [(sigma, nr_dims), ...])  (Lattice::init_params, Lattice.cu:107-132)."""
    block = cfg["lattice_gpu"]
    capacity = int(block["hash_table_capacity"])
    sigmas = []
    for i in range(int(block["nr_sigmas"])):
        toks = str(block[f"sigma_{i}"]).split()
        if len(toks) != 2:
            raise ValueError(f"sigma_{i} must be '<value> <nr_dims>', got {block[f'sigma_{i}']!r}")
        sigmas.append((float(toks[0]), int(float(toks[1]))))
    return capacity, sigmas


class _Params:
    _block = ""
    _defaults = {}

    def __init__(self, values=None):
        self._v = dict(self._defaults)
        if values:
            self._v.update(values)

    @classmethod
    def create(cls, cfg_or_path):
        cfg = cfg_or_path if isinstance(cfg_or_path, dict) else parse_cfg(cfg_or_path)
        return cls(cfg.get(cls._block, {}))

    def _get(self, key):
        return self._v[key]


def _getter(key):
    def get(self):
        return self._v[key]
    get.__name__ = key
    return get


class TrainParams(_Params):
    _block = "train"
    _defaults = dict(dataset_name="synthetic", with_viewer=False, with_visdom=False, with_tensorboard=False,
                     lr=1e-3, weight_decay=3e-4, save_checkpoint=False, checkpoint_path="")


for _k in TrainParams._defaults:
    setattr(TrainParams, _k, _getter(_k))


class EvalParams(_Params):
    _block = "eval"
    _defaults = dict(dataset_name="synthetic", with_viewer=False, checkpoint_path="", do_write_predictions=False,
                     output_predictions_path="")


for _k in EvalParams._defaults:
    setattr(EvalParams, _k, _getter(_k))


class ModelParams(_Params):
    _block = "model"
    # defaults = the ShapeNet architecture (config/lnn_train_shapenet.cfg:18-30)
    _defaults = dict(positions_mode="xyz", values_mode="none", pointnet_channels_per_layer=[16, 32, 64],
                     pointnet_start_nr_channels=32, nr_downsamples=3, nr_blocks_down_stage=[3, 3, 3],
                     nr_blocks_bottleneck=1, nr_blocks_up_stage=[2, 2, 2], nr_levels_down_with_normal_resnet=2,
                     nr_levels_up_with_normal_resnet=2, compression_factor=1.0, dropout_last_layer=0.0)


    def __init__(self, values=None):
        values = dict(values or {})
        # older cfgs spell it `pointnet_layers` (config/lnn_train_scannet.cfg:25 vs src/ModelParams.cxx:40)
        if "pointnet_layers" in values and "pointnet_channels_per_layer" not in values:
            values["pointnet_channels_per_layer"] = values.pop("pointnet_layers")
        super().__init__(values)


for _k in ModelParams._defaults:
    setattr(ModelParams, _k, _getter(_k))
