"""Keras weight files for the models of ``bpreveal_b200.models`` -- without h5py.

What the reference does with them: ``model.save(prefix + ".keras")`` (trainSoloModel.py:187,
trainCombinedModel.py:97-98, trainTransformationModel.py), ``ModelCheckpoint(<prefix>.checkpoint.keras)``
(callbacks.py:759-764) and ``utils.loadModel`` (utils.py:24-85: ``keras.models.load_model`` for
``.keras``; ``tf_keras`` for the pre-5.0 ``.model`` / ``.h5`` layouts).

A ``.keras`` file is a zip of ``config.json`` (the functional graph), ``metadata.json`` and
``model.weights.h5``.  Inside the HDF5 part Keras 3 addresses a layer by the snake-cased CLASS name
with a per-container counter in ``model.layers`` order -- ``layers/conv1d``, ``layers/conv1d_1``,
..., ``layers/dense``, nested models under ``layers/functional/layers/...`` -- and its variables as
``vars/0`` (kernel), ``vars/1`` (bias).  Legacy whole-model ``.h5`` files keep
``model_weights/<layer>/<layer>/{kernel:0,bias:0}`` plus a ``model_config`` attribute.

Reading is driven by the file's own ``config.json`` / ``model_config``: the architecture
(``soloModel`` / ``transformationModel`` / ``combinedModel`` arguments) is derived from the layer
names and shapes the reference's builders assign (models.py:39-48, 85-104, 141-172, 345-380), so
a file written by Keras is accepted, not only files written here.  Variable lookup tries the
class-counter path first and falls back to any group named after the layer (Keras 2 / tf_keras
layouts, and Keras 3 builds that may name by layer).  HDF5 access goes through ``h5lite``.

Honest status: the layouts above are restated from the published Keras 3 saving code
(``keras/src/saving/saving_lib.py``: ``_save_container_state`` / ``H5IOStore``), not checked
against a Keras-written file -- none exists offline (uncertainty 3).
``tests/test_keras_io_cpu.py`` builds such files from that description with an independent
writer path and loads them; when ``keras`` + ``h5py`` import on a box, ``tests/test_tf_probe.py``
round-trips a real file.
"""
from __future__ import annotations

import datetime
import io
import json
import re
import zipfile

import numpy as np

from . import h5lite

KERAS_VERSION = "3.5.0"  # the version whose layout is restated here


def to_snake_case(name):
    """keras/src/utils/naming.py ``to_snake_case``."""
    name = re.sub(r"\W+", "", name)
    name = re.sub("(.)([A-Z][a-z]+)", r"\1_\2", name)
    name = re.sub("([a-z])([A-Z])", r"\1_\2", name).lower()
    return name


# ---------------------------------------------------------------------------------------------
# architecture description  <->  Keras functional config
# ---------------------------------------------------------------------------------------------
_DT = {"module": "keras", "class_name": "DTypePolicy", "config": {"name": "float32"},
       "registered_name": None}


def _kt(shape, layer, node=0, idx=0):
    return {"class_name": "__keras_tensor__",
            "config": {"shape": list(shape), "dtype": "float32",
                       "keras_history": [layer, node, idx]}}


def _layer(cls, name, cfg, inputs, in_shape, module="keras.layers", registered=None, kwargs=None):
    """One entry of a functional config's ``layers`` list.  ``inputs``: list of
    (shape, producing layer, node, tensor index); a layer with several inputs takes them as one
    list argument (Add) unless ``kwargs`` says otherwise."""
    full = {"name": name, "trainable": True, "dtype": _DT}
    full.update(cfg)
    args = [_kt(*i) for i in inputs]
    if cls == "Add":
        args = [args]
    entry = {"module": module, "class_name": cls, "config": full, "registered_name": registered,
             "name": name,
             "inbound_nodes": [{"args": args, "kwargs": kwargs or {}}] if inputs else []}
    if in_shape is not None:
        entry["build_config"] = {"input_shape": in_shape}
    return entry


def _conv_cfg(filters, width, dilation, activation, ones=False):
    init = {"module": "keras.initializers", "class_name": "Ones" if ones else "GlorotUniform",
            "config": {} if ones else {"seed": None}, "registered_name": None}
    return {"filters": int(filters), "kernel_size": [int(width)], "strides": [1],
            "padding": "valid", "data_format": "channels_last", "dilation_rate": [int(dilation)],
            "groups": 1, "activation": activation, "use_bias": True,
            "kernel_initializer": init,
            "bias_initializer": {"module": "keras.initializers", "class_name": "Zeros",
                                 "config": {}, "registered_name": None},
            "kernel_regularizer": None, "bias_regularizer": None, "activity_regularizer": None,
            "kernel_constraint": None, "bias_constraint": None}


def _solo_layers(a, in_tensor=None):
    """Layer entries of ``soloModel`` (models.py:85-113).  Returns (layers, profile outputs,
    counts outputs, input entry name); outputs as (shape, layer, node, index)."""
    mn, Lin, F = a["modelName"], a["inputLength"], a["numFilters"]
    layers = []
    if in_tensor is None:
        layers.append({"module": "keras.layers", "class_name": "InputLayer",
                       "config": {"batch_shape": [None, Lin, 4], "dtype": "float32",
                                  "sparse": False, "name": f"{mn}_input"},
                       "registered_name": None, "name": f"{mn}_input", "inbound_nodes": []})
        cur = ([None, Lin, 4], f"{mn}_input", 0, 0)
    else:
        cur = in_tensor
    L = Lin - a["inputFilterWidth"] + 1
    layers.append(_layer("Conv1D", f"{mn}_initial_conv",
                         _conv_cfg(F, a["inputFilterWidth"], 1, "relu"), [cur], cur[0]))
    prev = ([None, L, F], f"{mn}_initial_conv", 0, 0)
    for i in range(a["numLayers"]):
        d = 2 ** (i + 1)
        Ln = L - 2 * d
        layers.append(_layer("Conv1D", f"{mn}_conv_{i}", _conv_cfg(F, 3, d, "relu"), [prev],
                             prev[0]))
        layers.append(_layer("Cropping1D", f"{mn}_crop{i}", {"cropping": [d, d]}, [prev],
                             prev[0]))
        conv = ([None, Ln, F], f"{mn}_conv_{i}", 0, 0)
        crop = ([None, Ln, F], f"{mn}_crop{i}", 0, 0)
        layers.append(_layer("Add", f"{mn}_add{i}", {}, [conv, crop], [conv[0], crop[0]]))
        prev, L = ([None, Ln, F], f"{mn}_add{i}", 0, 0), Ln
    profs, cnts = [], []
    Lout = L - a["outputFilterWidth"] + 1
    for h in a["heads"]:
        n, t = h["head-name"], int(h["num-tasks"])
        layers.append(_layer("Conv1D", f"solo_profile_{n}",
                             _conv_cfg(t, a["outputFilterWidth"], 1, "linear"), [prev], prev[0]))
        layers.append(_layer("GlobalAveragePooling1D", f"solo_counts_gap_{n}",
                             {"data_format": "channels_last", "keepdims": False}, [prev],
                             prev[0]))
        dense = {"units": 1, "activation": "linear", "use_bias": True,
                 "kernel_initializer": {"module": "keras.initializers",
                                        "class_name": "GlorotUniform", "config": {"seed": None},
                                        "registered_name": None},
                 "bias_initializer": {"module": "keras.initializers", "class_name": "Zeros",
                                      "config": {}, "registered_name": None},
                 "kernel_regularizer": None, "bias_regularizer": None,
                 "kernel_constraint": None, "bias_constraint": None}
        layers.append(_layer("Dense", f"solo_logcounts_{n}", dense,
                             [([None, F], f"solo_counts_gap_{n}", 0, 0)], [None, F]))
        profs.append(([None, Lout, t], f"solo_profile_{n}", 0, 0))
        cnts.append(([None, 1], f"solo_logcounts_{n}", 0, 0))
    return layers, profs, cnts


def _regression(name, src, shape):
    """layers.linearRegression (layers.py:35-46): Reshape -> Conv1D(1, 1, Ones) -> Reshape."""
    n = int(np.prod([s for s in shape[1:]]))
    flat = [None, n, 1]
    return [
        _layer("Reshape", name + "_r1", {"target_shape": [-1, 1]}, [src], src[0]),
        _layer("Conv1D", name + "_conv", _conv_cfg(1, 1, 1, "linear", ones=True),
               [(flat, name + "_r1", 0, 0)], flat),
        _layer("Reshape", name + "_r2", {"target_shape": list(shape[1:])},
               [(flat, name + "_conv", 0, 0)], flat),
    ], (list(shape), name + "_r2", 0, 0)


def _transform_layers(spec, tag, src):
    """models.py:118-175 for one output tensor."""
    if spec["name"] == "passthrough":
        return [], src
    layers, outs = [], []
    for t in spec["types"]:
        if t == "linear":
            ls, o = _regression(f"regress_linear_{tag}", src, src[0])
            layers += ls
        else:
            ls, o1 = _regression(f"{t}_in_linear_{tag}", src, src[0])
            layers += ls
            layers.append(_layer("Activation", f"{t}_activation_{tag}", {"activation": t}, [o1],
                                 o1[0]))
            ls, o = _regression(f"{t}_out_linear_{tag}",
                                (src[0], f"{t}_activation_{tag}", 0, 0), src[0])
            layers += ls
        outs.append(o)
    if len(outs) > 1:
        layers.append(_layer("Add", f"regress_sum_{tag}", {}, outs, [o[0] for o in outs]))
        return layers, (src[0], f"regress_sum_{tag}", 0, 0)
    return layers, outs[0]


def _functional(name, layers, inputs, outputs, trainable=True):
    return {"module": "keras", "class_name": "Functional",
            "config": {"name": name, "trainable": trainable, "layers": layers,
                       "input_layers": [[i[1], i[2], i[3]] for i in inputs],
                       "output_layers": [[o[1], o[2], o[3]] for o in outputs]},
            "registered_name": "Functional", "build_config": {"input_shape": None}}


def keras_config(arch):
    """Architecture description (``Model.get_config()``) -> Keras 3 functional config."""
    kind = arch["kind"]
    if kind == "solo":
        layers, profs, cnts = _solo_layers(arch)
        inp = ([None, arch["inputLength"], 4], f"{arch['modelName']}_input", 0, 0)
        return _functional(f"{arch['modelName']}_model", layers, [inp], profs + cnts)
    if kind == "transformation":
        solo = arch["solo"]
        layers, profs, cnts = _solo_layers(solo)
        for e in layers:
            e["config"]["trainable"] = False  # models.py:249
        po, co = [], []
        for h, p, c in zip(solo["heads"], profs, cnts):
            ls, o = _transform_layers(arch["profile"], f"profile_{h['head-name']}", p)
            layers += ls
            po.append(o)
            ls, o = _transform_layers(arch["counts"], f"logcounts_{h['head-name']}", c)
            layers += ls
            co.append(o)
        inp = ([None, solo["inputLength"], 4], f"{solo['modelName']}_input", 0, 0)
        return _functional("transformation_model", layers, [inp], po + co)
    if kind == "combined":
        res, bias = arch["residual"], arch["bias"]
        layers, profs, cnts = _solo_layers(res)
        inp = ([None, res["inputLength"], 4], f"{res['modelName']}_input", 0, 0)
        bcfg = keras_config(bias)
        bcfg["config"]["trainable"] = False
        bias_in = bias["solo"]["inputLength"] if bias["kind"] == "transformation" else \
            bias["inputLength"]
        flank = (res["inputLength"] - bias_in) // 2
        layers.append(_layer("Cropping1D", "auto_crop_input_tensor", {"cropping": [flank, flank]},
                             [inp], inp[0]))
        nested = dict(bcfg)
        nested["name"] = bcfg["config"]["name"]
        nested["inbound_nodes"] = [{"args": [[_kt([None, bias_in, 4], "auto_crop_input_tensor")]],
                                    "kwargs": {}}]
        layers.append(nested)
        H = len(res["heads"])
        po, co = [], []
        for i, (h, use) in enumerate(zip(res["heads"], arch["use-bias-counts"])):
            n, t = h["head-name"], int(h["num-tasks"])
            shp = [None, res["outputLength"], t]
            layers.append(_layer("Add", f"combined_add_profile_{n}", {},
                                 [(shp, nested["name"], 1, i), profs[i]], [shp, shp]))
            po.append((shp, f"combined_add_profile_{n}", 0, 0))
            if use:
                e = _layer("CountsLogSumExp", f"combined_logcounts_{n}", {},
                           [([None, 1], nested["name"], 1, H + i), cnts[i]], None,
                           module="bpreveal.layers", registered="CountsLogSumExp")
            else:
                e = _layer("Identity", f"combined_logcounts_{n}", {}, [cnts[i]], cnts[i][0])
            layers.append(e)
            co.append(([None, 1], f"combined_logcounts_{n}", 0, 0))
        return _functional("combined_model", layers, [inp], po + co)
    raise ValueError(f"unknown model kind {kind!r}")


def _layer_list(cfg):
    """``config.json`` / ``model_config`` -> the functional model's config dict."""
    if "config" in cfg and "layers" in cfg["config"]:
        return cfg["config"]
    raise ValueError("not a Keras functional-model configuration")


def _ksize(c):
    k = c["kernel_size"]
    return int(k[0] if isinstance(k, (list, tuple)) else k)


def arch_from_keras_config(cfg):
    """Keras functional config (as written by Keras 2 / 3 for the reference's builders, or by
    ``keras_config``) -> architecture description."""
    if cfg.get("module") == "bpreveal_b200.models":  # files of round 1
        return cfg["config"]
    fc = _layer_list(cfg)
    layers = fc["layers"]
    by_name = {e.get("name", e["config"].get("name")): e for e in layers}
    names = list(by_name)
    model_name = fc.get("name", "")

    def solo_arch(prefix_model):
        mn = prefix_model
        c0 = by_name[f"{mn}_initial_conv"]["config"]
        nl = 0
        while f"{mn}_conv_{nl}" in by_name:
            nl += 1
        inp = by_name.get(f"{mn}_input")
        shape = None
        if inp is not None:
            ic = inp["config"]
            shape = ic.get("batch_shape") or ic.get("batch_input_shape")
        # heads in output order (profiles first: models.py:112-114)
        heads, wout = [], None
        order = [o[0] for o in fc.get("output_layers", [])]
        prof_layers = [n for n in names if n.startswith("solo_profile_")]
        hn_by_pos = {}
        for pos, lname in enumerate(order):
            for tag in ("combined_add_profile_", "solo_profile_", "regress_sum_profile_"):
                if lname.startswith(tag):
                    hn_by_pos[pos] = lname[len(tag):]
            m = re.fullmatch(r"(?:regress|sigmoid_out|relu_out)_linear_profile_(.*)_r2", lname)
            if m:
                hn_by_pos[pos] = m.group(1)
        head_names = [hn_by_pos[p] for p in sorted(hn_by_pos)] or \
            [n[len("solo_profile_"):] for n in prof_layers]
        for hn in head_names:
            pc = by_name[f"solo_profile_{hn}"]["config"]
            heads.append({"head-name": hn, "num-tasks": int(pc["filters"])})
            wout = _ksize(pc)
        win = _ksize(c0)
        lin = int(shape[1]) if shape else None
        diff = (win - 1) + sum(2 * 2 ** (i + 1) for i in range(nl)) + (wout - 1)
        return {"kind": "solo", "modelName": mn, "inputLength": lin,
                "outputLength": lin - diff if lin is not None else None,
                "numFilters": int(c0["filters"]), "numLayers": nl, "inputFilterWidth": win,
                "outputFilterWidth": wout, "heads": heads}

    def spec_for(kind, heads):
        types = []
        hn = heads[0]["head-name"]
        if f"regress_linear_{kind}_{hn}_conv" in by_name:
            types.append("linear")
        for t in ("sigmoid", "relu"):
            if f"{t}_in_linear_{kind}_{hn}_conv" in by_name:
                types.append(t)
        return {"name": "simple", "types": types} if types else {"name": "passthrough"}

    init = [n for n in names if n.endswith("_initial_conv")]
    if model_name == "combined_model" or any(n.startswith("combined_add_profile_") for n in names):
        res = solo_arch("residual")
        nested = next(e for e in layers if e.get("class_name") in ("Functional", "Model")
                      and "layers" in e.get("config", {}))
        bias = arch_from_keras_config(nested)
        use = []
        for h in res["heads"]:
            e = by_name[f"combined_logcounts_{h['head-name']}"]
            use.append(e["class_name"] == "CountsLogSumExp")
        return {"kind": "combined", "residual": res, "bias": bias, "use-bias-counts": use}
    if model_name == "transformation_model" or any("_linear_profile_" in n or
                                                   "_linear_logcounts_" in n for n in names):
        mn = init[0][:-len("_initial_conv")]
        solo = solo_arch(mn)
        return {"kind": "transformation", "solo": solo,
                "profile": spec_for("profile", solo["heads"]),
                "counts": spec_for("logcounts", solo["heads"])}
    if not init:
        raise ValueError("no '<name>_initial_conv' layer: not a BPReveal model")
    return solo_arch(init[0][:-len("_initial_conv")])


def weight_paths(cfg, prefix=""):
    """Functional config -> {scoped layer name: 'layers/<snake class>[_k]' path} following
    ``saving_lib._save_container_state``: counters per container in ``layers`` order; nested
    functional models recurse with their name as the scope."""
    fc = _layer_list(cfg)
    used, out = {}, {}
    for e in fc["layers"]:
        cls = e["class_name"]
        base = to_snake_case(cls)
        if base in used:
            used[base] += 1
            key = f"{base}_{used[base]}"
        else:
            used[base] = 0
            key = base
        name = e.get("name", e["config"].get("name"))
        path = f"{prefix}layers/{key}"
        if "layers" in e.get("config", {}):
            sub = weight_paths(e, prefix=path + "/")
            out.update({f"{name}/{k}": v for k, v in sub.items()})
        else:
            out[name] = path
    return out


# ---------------------------------------------------------------------------------------------
# save
# ---------------------------------------------------------------------------------------------
def _split_named(named_weights):
    """[('scope/layer/kernel', array)] -> {scoped layer: {'kernel': a, 'bias': b}} (ordered)."""
    out = {}
    for name, arr in named_weights:
        layer, var = name.rsplit("/", 1)
        out.setdefault(layer, {})[var] = np.asarray(arr, dtype=np.float32)
    return out


def weights_h5_bytes(arch, named_weights):
    """``model.weights.h5`` in the Keras 3 layout."""
    cfg = keras_config(arch)
    paths = weight_paths(cfg)
    w = h5lite.Writer()
    for layer, vs in _split_named(named_weights).items():
        if layer not in paths:
            raise KeyError(f"layer {layer!r} is not part of the model configuration")
        for i, var in enumerate(("kernel", "bias")):
            w.create_dataset(f"{paths[layer]}/vars/{i}", data=vs[var])
    # Keras creates an (empty) vars group for every layer it visits
    for path in paths.values():
        w.create_group(path + "/vars")
    return w.tobytes()


def legacy_h5_bytes(arch, named_weights):
    """Whole-model legacy ``.h5`` (Keras 2 / tf_keras ``save_format='h5'``)."""
    cfg = keras_config(arch)
    w = h5lite.Writer()
    w.attrs("")["keras_version"] = KERAS_VERSION
    w.attrs("")["backend"] = "tensorflow"
    w.attrs("")["model_config"] = json.dumps(cfg)
    groups = _split_named(named_weights)
    w.create_group("model_weights")
    w.attrs("model_weights")["layer_names"] = [k.encode("utf-8") for k in groups]
    w.attrs("model_weights")["keras_version"] = KERAS_VERSION
    w.attrs("model_weights")["backend"] = "tensorflow"
    for layer, vs in groups.items():
        leaf = layer.rsplit("/", 1)[-1]
        g = f"model_weights/{layer}"
        w.create_group(g)
        w.attrs(g)["weight_names"] = [f"{leaf}/{v}:0".encode("utf-8") for v in ("kernel", "bias")]
        for var in ("kernel", "bias"):
            w.create_dataset(f"{g}/{leaf}/{var}:0", data=vs[var])
    return w.tobytes()


def save_keras(path, arch, named_weights):
    """``Model.save`` for ``.keras`` (zip) and ``.h5`` / ``.hdf5`` (legacy whole-model)."""
    if path.endswith((".h5", ".hdf5")):
        with open(path, "wb") as fp:
            fp.write(legacy_h5_bytes(arch, named_weights))
        return
    cfg = keras_config(arch)
    meta = {"keras_version": KERAS_VERSION, "writer": "bpreveal_b200",
            "date_saved": datetime.datetime.now().strftime("%Y-%m-%d@%H:%M:%S")}
    with zipfile.ZipFile(path, "w", zipfile.ZIP_STORED) as z:
        z.writestr("metadata.json", json.dumps(meta))
        z.writestr("config.json", json.dumps(cfg))
        z.writestr("model.weights.h5", weights_h5_bytes(arch, named_weights))


# ---------------------------------------------------------------------------------------------
# load
# ---------------------------------------------------------------------------------------------
def _collect_datasets(f):
    out = {}
    f.visititems(lambda name, obj: out.__setitem__(name, obj)
                 if isinstance(obj, h5lite.Dataset) else None)
    return out


def _find_vars(datasets, scoped_layer, cls_path):
    """kernel / bias of one layer: the class-counter path first, then anything named after the
    layer."""
    if cls_path is not None:
        k, b = datasets.get(f"{cls_path}/vars/0"), datasets.get(f"{cls_path}/vars/1")
        if k is not None and b is not None:
            return k[...], b[...]
    parts = scoped_layer.split("/")
    leaf = parts[-1]
    cand = {}
    for name, ds in datasets.items():
        comps = name.split("/")
        if leaf not in comps:
            continue
        if len(parts) > 1 and parts[0] not in comps:
            continue  # nested model: its name must be on the path
        if len(parts) == 1 and any(c in ("transformation_model", "functional") for c in comps) \
                and "residual" not in leaf and leaf.startswith("solo_"):
            # the top-level model's own solo_* layers, not the nested bias model's
            continue
        last = comps[-1].split(":")[0]
        var = {"0": "kernel", "1": "bias"}.get(last, last)
        if var in ("kernel", "bias"):
            cand.setdefault(var, ds)
    if "kernel" in cand and "bias" in cand:
        return cand["kernel"][...], cand["bias"][...]
    raise KeyError(f"no kernel / bias found for layer {scoped_layer!r} "
                   f"(tried {cls_path}/vars/{{0,1}} and name-based lookup)")


def load_keras(path):
    """-> (architecture description, {scoped 'layer/kernel' | 'layer/bias': array})."""
    if zipfile.is_zipfile(path):
        with zipfile.ZipFile(path) as z:
            cfg = json.loads(z.read("config.json"))
            raw = z.read("model.weights.h5")
        f = h5lite.File(raw)
    else:
        f = h5lite.File(path)
        mc = f.attrs.get("model_config")
        if mc is None:
            raise ValueError(f"{path}: an HDF5 file without a 'model_config' attribute holds "
                             "weights only; build the model and use load_weights_into()")
        cfg = json.loads(h5lite.as_str(mc))
    arch = arch_from_keras_config(cfg)
    if cfg.get("module") == "bpreveal_b200.models":
        paths = {}
    else:
        paths = weight_paths(cfg)
    return arch, read_weights(f, paths)


def read_weights(f, paths):
    """Every (kernel, bias) pair the file holds for the layers in ``paths`` (or, with an empty
    map, every ``<layer>/vars/{0,1}`` / ``<layer>/{kernel,bias}:0`` pair by name)."""
    datasets = _collect_datasets(f)
    weights = {}
    if paths:
        for layer, p in paths.items():
            try:
                k, b = _find_vars(datasets, layer, p)
            except KeyError:
                continue  # layers without variables
            weights[f"{layer}/kernel"], weights[f"{layer}/bias"] = k, b
        return weights
    for name, ds in datasets.items():
        comps = name.split("/")
        if len(comps) >= 3 and comps[-2] == "vars":
            var = {"0": "kernel", "1": "bias"}.get(comps[-1], comps[-1])
            weights[f"{comps[-3]}/{var}"] = ds[...]
        elif len(comps) >= 2:
            weights[f"{comps[-2]}/{comps[-1].split(':')[0]}"] = ds[...]
    return weights
