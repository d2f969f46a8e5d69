#!/usr/bin/env python
"""Generates the committed fixtures under tests/golden/.  Run in the build container only:

    python tests/golden/make_golden.py

Two files are produced:

* ``reference_answers.npz`` -- outputs of the REFERENCE'S OWN Python/C code, imported from
  /root/reference/pkg (read-only) with its heavyweight dependencies (tensorflow, keras, h5py,
  pysam, pybedtools, ...) replaced by inert stub modules.  Only functions that are pure
  numpy / plain C can be executed that way: lengthCalc.getLengthDifference, utils.oneHotEncode /
  oneHotDecode / logitsToProfile, interpretUtils._ShapBatcher.generateShuffles (kmer 1) and
  combineMultAndDiffref, shap.standard_combine_mult_and_diffref, and libslide.c (compiled into
  oracle/_ref by oracle/Makefile).  These pin the corresponding oracle functions.
* ``oracle_vectors.npz`` -- outputs of the oracle itself (fp64) on seeded inputs for the tensor
  arithmetic the reference delegates to TensorFlow/Keras (not runnable offline): forward, losses,
  gradients, five Keras-Adam steps, DeepSHAP (PISA and flat metrics), transformation and combined
  models.  These are regression pins, not reference pins ("parity unpinned").
"""
import ctypes
import importlib.abc
import importlib.machinery
import os
import sys
import types
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_PKG = "/root/reference/pkg"

STUB_ROOTS = ("tensorflow", "keras", "tf_keras", "tensorflow_probability", "h5py", "pysam",
              "pybedtools", "pyBigWig", "matplotlib", "tqdm")
STUB_EXACT = ("bpreveal.internal.libushuffle", "bpreveal.internal.libslide")


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return mock.MagicMock(name=f"{self.__name__}.{name}")


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in STUB_ROOTS or name in STUB_EXACT:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


def import_reference():
    sys.meta_path.insert(0, _StubFinder())
    sys.path.insert(0, REF_PKG)
    # numpy >= 2.3 removed the binary mode of np.fromstring which utils.oneHotEncode (utils.py:479)
    # still calls; restore exactly that behaviour.
    np.fromstring = lambda s, dtype: np.frombuffer(s.encode("ascii"), dtype=dtype)  # noqa: E731
    import bpreveal.lengthCalc as lengthCalc
    import bpreveal.utils as utils
    import bpreveal.internal.interpretUtils as interpretUtils
    import bpreveal.internal.shap as shap
    return lengthCalc, utils, interpretUtils, shap


def reference_answers():
    lengthCalc, utils, interpretUtils, shap = import_reference()
    out = {}
    # lengthCalc.py:71-123
    grid = [(L, wi, wo) for L in (0, 1, 3, 9, 11) for wi in (3, 7, 25) for wo in (3, 7, 23, 75)]
    out["lengthcalc_grid"] = np.array(grid, dtype=np.int64)
    out["lengthcalc_diff"] = np.array(
        [lengthCalc.getLengthDifference(L, [wi], wo, False) for L, wi, wo in grid], dtype=np.int64)
    # utils.py:420-486
    rng = np.random.default_rng(1)
    seqs = ["ACGTTT", "acgtNnACGT", "".join(rng.choice(list("ACGT"), size=200))]
    for i, s in enumerate(seqs):
        out[f"ohe_seq{i}"] = np.array(s)
        out[f"ohe_out{i}"] = np.asarray(utils.oneHotEncode(s, allowN=True))
    # utils.py:523-572
    lg = rng.standard_normal((50, 2)).astype(np.float32) * 3
    out["l2p_logits"] = lg
    out["l2p_logcounts"] = np.float32(4.25)
    out["l2p_profile"] = np.asarray(utils.logitsToProfile(lg, np.float32(4.25)))
    # interpretUtils.py:1216-1222 (kmer 1 branch; `self` only supplies kmerSize / numShuffles)
    x = np.asarray(utils.oneHotEncode(seqs[2]))
    fake = types.SimpleNamespace(kmerSize=1, numShuffles=5)
    shuf = interpretUtils._ShapBatcher.generateShuffles(fake, [x])[0]
    out["shuffle_x"] = x
    out["shuffle_refs"] = np.asarray(shuf)
    # combiners: shap.py:26-30, interpretUtils.py:1295-1312
    mult = rng.standard_normal(shuf.shape)
    out["comb_mult"] = mult
    out["comb_standard"] = np.asarray(
        shap.standard_combine_mult_and_diffref([mult], [x[None].astype(float)],
                                               [shuf.astype(float)])[0])
    out["comb_hypothetical"] = np.asarray(
        interpretUtils.combineMultAndDiffref([mult], [x], [shuf])[0])
    # libslide.c:44-91 through the library compiled from the reference sources
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libslide.so"))
    rows, scols, dcols, depth = 7, 40, 31, 3
    src = rng.standard_normal((rows, scols, depth)).astype(np.float32)
    srcc = rng.integers(0, 2, size=(rows, scols, 4)).astype(np.uint8)
    rowIdx = rng.permutation(rows).astype(np.int32)
    colIdx = rng.integers(0, scols - dcols + 1, size=rows).astype(np.int32)
    dst = np.zeros((rows, dcols, depth), dtype=np.float32)
    dstc = np.zeros((rows, dcols, 4), dtype=np.float32)  # slideChar converts u8 -> float
    P = ctypes.c_void_p
    lib.slide(P(src.ctypes.data), P(dst.ctypes.data), ctypes.c_int(rows), ctypes.c_int(scols),
              ctypes.c_int(dcols), ctypes.c_int(depth), P(rowIdx.ctypes.data),
              P(colIdx.ctypes.data))
    lib.slideChar(P(srcc.ctypes.data), P(dstc.ctypes.data), ctypes.c_int(rows),
                  ctypes.c_int(scols), ctypes.c_int(dcols), ctypes.c_int(4),
                  P(rowIdx.ctypes.data), P(colIdx.ctypes.data))
    out.update(slide_src=src, slide_srcc=srcc, slide_rows=rowIdx, slide_cols=colIdx,
               slide_dst=dst, slide_dstc=dstc)
    return out


def oracle_vectors():
    from oracle import bpnet_oracle as O
    out = {}
    cfg = dict(numFilters=8, numLayers=3, inputFilterWidth=3, outputFilterWidth=3,
               headTasks=[2, 1])
    Lin, Lout = 52, 20
    params = O.init_solo_params(seed=1234, biasScale=0.1, scale=1.5, **cfg)
    for k, v in params.items():
        out[f"p_{k}"] = v
    B = 4
    x = O.synthetic_sequences(B, Lin, seed=5)
    counts = O.synthetic_profiles(B, Lout, 3, seed=5, meanReads=30.0)
    out["x"], out["counts"] = x, counts
    p = O.to_torch(params, torch.float64, requires_grad=True)
    xt = torch.tensor(x, dtype=torch.float64)
    ct = torch.tensor(counts, dtype=torch.float64)
    profs, cnts = O.solo_forward(xt, p)
    trueP = [ct[:, :, 0:2], ct[:, :, 2:3]]
    trueC = [torch.log(t.sum(dim=(1, 2))) for t in trueP]
    pw, lam = [1.0, 1.5], [2.0, 3.0]
    tot, pl, cl = O.total_loss(profs, cnts, trueP, trueC, pw, lam)
    tot.backward()
    out["logits"] = torch.cat(profs, dim=2).detach().numpy()
    out["logcounts"] = torch.cat(cnts, dim=1).detach().numpy()
    out["losses"] = np.array([v.item() for v in pl] + [v.item() for v in cl] + [tot.item()])
    for k, v in p.items():
        out[f"g_{k}"] = v.grad.numpy().copy()
    # five Adam steps
    p = O.to_torch(params, torch.float64, requires_grad=True)
    m = {k: torch.zeros_like(v) for k, v in p.items()}
    v_ = {k: torch.zeros_like(v) for k, v in p.items()}
    curve = []
    for step in range(1, 6):
        profs, cnts = O.solo_forward(xt, p)
        tot, pl, cl = O.total_loss(profs, cnts, trueP, trueC, pw, lam)
        for t in p.values():
            t.grad = None
        tot.backward()
        with torch.no_grad():
            for k in p:
                O.adam_step(p[k], p[k].grad, m[k], v_[k], step, 0.004)
        curve.append(tot.item())
    out["adam_curve"] = np.array(curve)
    out["adam_conv1_w"] = p["conv1_w"].detach().numpy().copy()
    # DeepSHAP
    p64 = O.to_torch(params, torch.float64)
    refs = O.generate_shuffles(x[0], 6)
    phi, val, rvals = O.deepshap(x[0], refs, O.solo_shap_forward(p64),
                                 lambda pr, cn: O.pisa_metric(pr, cn, 0, 1))
    out["pisa_phi"], out["pisa_val"], out["pisa_refvals"] = phi, np.float64(val), rvals
    phi_h, _, _ = O.deepshap(x[0], refs, O.solo_shap_forward(p64),
                             lambda pr, cn: O.profile_metric(pr, cn, 0, [0, 1], shap=True),
                             hypothetical=True)
    out["flat_profile_hyp"] = phi_h
    phi_c, _, _ = O.deepshap(x[0], refs, O.solo_shap_forward(p64),
                             lambda pr, cn: O.counts_metric(pr, cn, 1), hypothetical=True)
    out["flat_counts_hyp"] = phi_c
    # transformation + combined (bias model = same architecture, shorter input)
    spec_p = {"name": "simple", "types": ["linear", "sigmoid"]}
    spec_c = {"name": "simple", "types": ["linear"]}
    tp = O.init_transform_params(2, spec_p, spec_c)
    rng = np.random.default_rng(3)
    for k in tp:
        tp[k] = (tp[k] + rng.standard_normal(2) * 0.2).astype(np.float32)
        out[f"t_{k}"] = tp[k]
    tpt = {k: torch.tensor(v, dtype=torch.float64) for k, v in tp.items()}
    rparams = O.init_solo_params(numFilters=8, numLayers=4, inputFilterWidth=3,
                                 outputFilterWidth=3, headTasks=[2, 1], seed=77, biasScale=0.1)
    for k, v in rparams.items():
        out[f"r_{k}"] = v
    LinR = Lout + O.length_difference(4, 3, 3)
    xr = O.synthetic_sequences(B, LinR, seed=9)
    out["xr"] = xr
    profs, cnts, _ = O.combined_forward(torch.tensor(xr, dtype=torch.float64),
                                        O.to_torch(rparams), p64, tpt, spec_p, spec_c,
                                        [True, False], Lin)
    out["comb_logits"] = torch.cat(profs, dim=2).numpy()
    out["comb_logcounts"] = torch.cat(cnts, dim=1).numpy()
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **oracle_vectors())
    np.savez_compressed(os.path.join(HERE, "reference_answers.npz"), **reference_answers())
    print("wrote", os.listdir(HERE))
