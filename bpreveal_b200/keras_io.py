"""Keras-3 ``.keras`` / legacy ``.h5`` weight files for the models of ``bpreveal_b200.models``.

A ``.keras`` file is a zip of ``config.json``, ``metadata.json`` and ``model.weights.h5``; the HDF5
part needs ``h5py``, which the build image does not ship ("hard parts").  The code
below is therefore gated on ``import h5py`` and fails loudly without it; the always-available
native format is the ``.npz`` written by ``Model.save`` (same variable names).

Variable paths follow the layer names of the reference graph (models.py:39-48, 85-104):
``layers/<layer name>/vars/{0: kernel, 1: bias}`` -- unverified against a real Keras file (none is
available offline; SURVEY.md 9, uncertainty 3), so ``load_keras`` also accepts any group whose
last path component is the layer name.
"""
from __future__ import annotations

import io
import json
import zipfile

import numpy as np


def _h5py():
    try:
        import h5py  # noqa: PLC0415
    except ImportError as e:  # pragma: no cover - depends on the image
        raise RuntimeError(
            "reading / writing .keras and .h5 files needs h5py, which is not installed here; "
            "use the native .npz format (Model.save('x.npz') / models.loadModel('x.npz'))") from e
    return h5py


def save_keras(path, arch, named_weights):
    h5py = _h5py()
    buf = io.BytesIO()
    with h5py.File(buf, "w") as f:
        for name, arr in named_weights:
            layer, var = name.rsplit("/", 1)
            idx = {"kernel": "0", "bias": "1"}[var]
            f.create_dataset(f"layers/{layer}/vars/{idx}", data=np.asarray(arr, dtype=np.float32))
    if path.endswith(".h5"):
        with open(path, "wb") as fp:
            fp.write(buf.getvalue())
        return
    with zipfile.ZipFile(path, "w") as z:
        z.writestr("config.json", json.dumps({"module": "bpreveal_b200.models", "config": arch}))
        z.writestr("metadata.json", json.dumps({"keras_version": "3", "writer": "bpreveal_b200"}))
        z.writestr("model.weights.h5", buf.getvalue())


def load_keras(path):
    h5py = _h5py()
    arch = None
    if zipfile.is_zipfile(path):
        with zipfile.ZipFile(path) as z:
            cfg = json.loads(z.read("config.json"))
            arch = cfg.get("config") if cfg.get("module") == "bpreveal_b200.models" else None
            raw = z.read("model.weights.h5")
        f = h5py.File(io.BytesIO(raw), "r")
    else:
        f = h5py.File(path, "r")
    weights = {}

    def visit(name, obj):
        if isinstance(obj, h5py.Dataset):
            parts = name.split("/")
            if len(parts) >= 3 and parts[-2] == "vars":
                var = {"0": "kernel", "1": "bias"}.get(parts[-1], parts[-1])
                weights[f"{parts[-3]}/{var}"] = np.asarray(obj)
            else:  # legacy: <layer>/<layer>/kernel:0
                weights[f"{parts[-2]}/{parts[-1].split(':')[0]}"] = np.asarray(obj)
    f.visititems(visit)
    if arch is None:
        raise ValueError(f"{path} does not carry a bpreveal_b200 architecture description; build "
                         "the model with models.soloModel(...) and call set_weights(dict)")
    return arch, weights
