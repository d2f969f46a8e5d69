"""Keras weight files without h5py (keras_io + h5lite): the architecture is recovered from the
file's own layer configuration, variables are found by Keras 3's class-counter paths
(saving_lib._save_container_state) or by layer name (legacy layouts)."""
import io
import json
import zipfile

import numpy as np
import pytest

from bpreveal_b200 import h5lite, keras_io

HEADS = [{"head-name": "oct4", "num-tasks": 2}, {"head-name": "sox2", "num-tasks": 1}]
SOLO = {"kind": "solo", "modelName": "solo", "inputLength": 52 + 2, "outputLength": 20,
        "numFilters": 8, "numLayers": 3, "inputFilterWidth": 5, "outputFilterWidth": 3,
        "heads": HEADS}
TRANS = {"kind": "transformation", "solo": SOLO,
         "profile": {"name": "simple", "types": ["linear", "sigmoid"]},
         "counts": {"name": "simple", "types": ["linear"]}}
RESID = dict(SOLO, modelName="residual", inputLength=54 + 16 + 2, numLayers=4, numFilters=16,
             inputFilterWidth=7, outputFilterWidth=5, outputLength=20 - 4 + 2 + 0)


def solo_named(arch, rng, scope=""):
    F, L = arch["numFilters"], arch["numLayers"]
    mn = arch["modelName"]
    out = [(f"{scope}{mn}_initial_conv/kernel", rng.standard_normal((arch["inputFilterWidth"], 4, F))),
           (f"{scope}{mn}_initial_conv/bias", rng.standard_normal(F))]
    for i in range(L):
        out += [(f"{scope}{mn}_conv_{i}/kernel", rng.standard_normal((3, F, F))),
                (f"{scope}{mn}_conv_{i}/bias", rng.standard_normal(F))]
    for h in arch["heads"]:
        n, t = h["head-name"], h["num-tasks"]
        out += [(f"{scope}solo_profile_{n}/kernel", rng.standard_normal((arch["outputFilterWidth"], F, t))),
                (f"{scope}solo_profile_{n}/bias", rng.standard_normal(t)),
                (f"{scope}solo_logcounts_{n}/kernel", rng.standard_normal((F, 1))),
                (f"{scope}solo_logcounts_{n}/bias", rng.standard_normal(1))]
    return [(k, v.astype(np.float32)) for k, v in out]


def trans_named(arch, rng, scope=""):
    out = solo_named(arch["solo"], rng, scope)
    for h in arch["solo"]["heads"]:
        n = h["head-name"]
        for tag, spec in (("profile", arch["profile"]), ("logcounts", arch["counts"])):
            for t in spec.get("types", []):
                layers = [f"regress_linear_{tag}_{n}_conv"] if t == "linear" else \
                    [f"{t}_in_linear_{tag}_{n}_conv", f"{t}_out_linear_{tag}_{n}_conv"]
                for ln in layers:
                    out += [(f"{scope}{ln}/kernel", rng.standard_normal((1, 1, 1)).astype(np.float32)),
                            (f"{scope}{ln}/bias", rng.standard_normal(1).astype(np.float32))]
    return out


def fix_lengths(a):
    a = dict(a)
    diff = (a["inputFilterWidth"] - 1) + sum(2 * 2 ** (i + 1) for i in range(a["numLayers"])) + \
        (a["outputFilterWidth"] - 1)
    a["inputLength"] = a["outputLength"] + diff
    return a


SOLO = fix_lengths(SOLO)
TRANS["solo"] = SOLO
RESID = fix_lengths(dict(RESID, outputLength=20))
COMB = {"kind": "combined", "residual": RESID, "bias": TRANS, "use-bias-counts": [True, False]}


def test_snake_case_matches_keras_naming():
    for cls, want in (("Conv1D", "conv1d"), ("InputLayer", "input_layer"), ("Dense", "dense"),
                      ("GlobalAveragePooling1D", "global_average_pooling1d"),
                      ("Cropping1D", "cropping1d"), ("CountsLogSumExp", "counts_log_sum_exp"),
                      ("Functional", "functional"), ("Add", "add")):
        assert keras_io.to_snake_case(cls) == want


def test_config_round_trip_recovers_every_architecture():
    for arch in (SOLO, TRANS, COMB):
        cfg = keras_io.keras_config(arch)
        # what Keras would hand back after json round trip
        back = keras_io.arch_from_keras_config(json.loads(json.dumps(cfg)))
        assert back == arch, (arch["kind"], back)
    cfg = keras_io.keras_config(SOLO)
    names = [e["name"] for e in cfg["config"]["layers"]]
    # the reference's layer names and order (models.py:85-113)
    assert names[:6] == ["solo_input", "solo_initial_conv", "solo_conv_0", "solo_crop0",
                         "solo_add0", "solo_conv_1"]
    assert cfg["config"]["output_layers"] == [["solo_profile_oct4", 0, 0], ["solo_profile_sox2", 0, 0],
                                              ["solo_logcounts_oct4", 0, 0],
                                              ["solo_logcounts_sox2", 0, 0]]
    conv1 = cfg["config"]["layers"][3 - 1]
    assert conv1["config"]["dilation_rate"] == [2] and conv1["config"]["activation"] == "relu"
    assert conv1["inbound_nodes"][0]["args"][0]["config"]["keras_history"] == ["solo_initial_conv", 0, 0]


def test_weight_paths_follow_class_counters():
    paths = keras_io.weight_paths(keras_io.keras_config(SOLO))
    assert paths["solo_initial_conv"] == "layers/conv1d"
    assert paths["solo_conv_0"] == "layers/conv1d_1" and paths["solo_conv_2"] == "layers/conv1d_3"
    assert paths["solo_profile_oct4"] == "layers/conv1d_4"
    assert paths["solo_logcounts_oct4"] == "layers/dense"
    assert paths["solo_logcounts_sox2"] == "layers/dense_1"
    cp = keras_io.weight_paths(keras_io.keras_config(COMB))
    assert cp["transformation_model/solo_initial_conv"] == "layers/functional/layers/conv1d"
    assert cp["residual_initial_conv"] == "layers/conv1d"
    assert cp["solo_profile_oct4"].startswith("layers/conv1d_")


@pytest.mark.parametrize("ext", [".keras", ".h5"])
@pytest.mark.parametrize("kind", ["solo", "transformation", "combined"])
def test_save_load_round_trip(tmp_path, ext, kind):
    rng = np.random.default_rng(3)
    if kind == "solo":
        arch, named = SOLO, solo_named(SOLO, rng)
    elif kind == "transformation":
        arch, named = TRANS, trans_named(TRANS, rng)
    else:
        arch = COMB
        named = solo_named(RESID, rng) + trans_named(TRANS, rng, scope="transformation_model/")
    path = str(tmp_path / ("m" + ext))
    keras_io.save_keras(path, arch, named)
    arch2, weights = keras_io.load_keras(path)
    assert arch2 == arch
    assert set(weights) == {k for k, _ in named}
    for k, v in named:
        assert np.array_equal(weights[k], v), k
    if ext == ".keras":
        with zipfile.ZipFile(path) as z:
            assert sorted(z.namelist()) == ["config.json", "metadata.json", "model.weights.h5"]
            assert json.loads(z.read("metadata.json"))["keras_version"].startswith("3.")


def test_foreign_keras3_file_built_from_the_published_layout(tmp_path):
    """A ``.keras`` file assembled the way Keras 3 does it -- Keras-style config (``batch_shape``,
    per-layer ``module`` / ``registered_name``), variables under class-counter groups, an
    ``optimizer`` subtree, weights of layers this code knows nothing about elsewhere in the tree
    -- not produced by keras_io.save_keras: the loader must not depend on its own writer."""
    rng = np.random.default_rng(5)
    named = dict(solo_named(SOLO, rng))
    layers = [{"module": "keras.layers", "class_name": "InputLayer", "registered_name": None,
               "name": "solo_input", "inbound_nodes": [],
               "config": {"batch_shape": [None, SOLO["inputLength"], 4], "dtype": "float32",
                          "sparse": False, "name": "solo_input"}}]

    def conv(name, filters, k, d):
        return {"module": "keras.layers", "class_name": "Conv1D", "registered_name": None,
                "name": name, "inbound_nodes": [],
                "config": {"name": name, "filters": filters, "kernel_size": [k],
                           "dilation_rate": [d], "padding": "valid", "activation": "relu"}}

    def other(cls, name):
        return {"module": "keras.layers", "class_name": cls, "registered_name": None, "name": name,
                "inbound_nodes": [], "config": {"name": name}}

    layers.append(conv("solo_initial_conv", 8, 5, 1))
    for i in range(3):
        layers += [conv(f"solo_conv_{i}", 8, 3, 2 ** (i + 1)), other("Cropping1D", f"solo_crop{i}"),
                   other("Add", f"solo_add{i}")]
    for h in HEADS:
        n = h["head-name"]
        layers += [conv(f"solo_profile_{n}", h["num-tasks"], 3, 1),
                   other("GlobalAveragePooling1D", f"solo_counts_gap_{n}"),
                   {"module": "keras.layers", "class_name": "Dense", "registered_name": None,
                    "name": f"solo_logcounts_{n}", "inbound_nodes": [],
                    "config": {"name": f"solo_logcounts_{n}", "units": 1}}]
    cfg = {"module": "keras", "class_name": "Functional", "registered_name": "Functional",
           "config": {"name": "solo_model", "trainable": True, "layers": layers,
                      "input_layers": [["solo_input", 0, 0]],
                      "output_layers": [[f"solo_profile_{h['head-name']}", 0, 0] for h in HEADS] +
                                       [[f"solo_logcounts_{h['head-name']}", 0, 0] for h in HEADS]},
           "compile_config": {"optimizer": "adam"}}
    w = h5lite.Writer()
    conv_i = dense_i = 0
    for e in layers:
        if e["class_name"] == "Conv1D":
            g = "layers/conv1d" + (f"_{conv_i}" if conv_i else "")
            conv_i += 1
        elif e["class_name"] == "Dense":
            g = "layers/dense" + (f"_{dense_i}" if dense_i else "")
            dense_i += 1
        else:
            continue
        w.create_dataset(g + "/vars/0", data=named[e["name"] + "/kernel"])
        w.create_dataset(g + "/vars/1", data=named[e["name"] + "/bias"])
    w.create_dataset("optimizer/vars/0", data=np.int64(7))           # iteration counter
    w.create_dataset("optimizer/vars/1", data=np.float32(0.004))
    buf = io.BytesIO()
    with zipfile.ZipFile(buf, "w") as z:
        z.writestr("metadata.json", json.dumps({"keras_version": "3.6.0", "date_saved": "x"}))
        z.writestr("config.json", json.dumps(cfg))
        z.writestr("model.weights.h5", w.tobytes())
    path = tmp_path / "foreign.keras"
    path.write_bytes(buf.getvalue())
    arch, weights = keras_io.load_keras(str(path))
    assert arch == SOLO
    for k, v in named.items():
        assert np.array_equal(weights[k], v), k


def test_legacy_h5_with_name_based_groups(tmp_path):
    """Keras 2 / tf_keras whole-model h5: ``model_config`` attribute with ``batch_input_shape``,
    variables at ``model_weights/<layer>/<layer>/kernel:0``."""
    rng = np.random.default_rng(6)
    named = solo_named(SOLO, rng)
    cfg = keras_io.keras_config(SOLO)
    inp = cfg["config"]["layers"][0]["config"]
    inp["batch_input_shape"] = inp.pop("batch_shape")
    for e in cfg["config"]["layers"]:
        if "kernel_size" in e["config"]:
            e["config"]["kernel_size"] = tuple(e["config"]["kernel_size"])
    w = h5lite.Writer()
    w.attrs("")["model_config"] = json.dumps(cfg)
    for k, v in named:
        layer, var = k.split("/")
        w.create_dataset(f"model_weights/{layer}/{layer}/{var}:0", data=v)
    p = tmp_path / "old.h5"
    p.write_bytes(w.tobytes())
    arch, weights = keras_io.load_keras(str(p))
    assert arch == SOLO
    for k, v in named:
        assert np.array_equal(weights[k], v), k
    # weights-only files are refused with an explanation
    w2 = h5lite.Writer()
    w2.create_dataset("layers/conv1d/vars/0", data=np.zeros(3, dtype=np.float32))
    q = tmp_path / "w.h5"
    q.write_bytes(w2.tobytes())
    with pytest.raises(ValueError, match="model_config"):
        keras_io.load_keras(str(q))
