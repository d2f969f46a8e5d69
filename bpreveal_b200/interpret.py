"""DeepSHAP attribution on the device (interpretPisa / interpretFlat hot loop).

Replaces the reference's ``_ShapBatcher.runPrediction`` + ``TFDeepExplainer.shap_values``
(src/internal/interpretUtils.py:1242-1292, src/internal/shap.py:347-394): per query the reference
runs 21 + 40 forward and 40 backward passes through TensorFlow; here the query is run once, the n
references once, and the rescale-rule backward once per (query, reference) pair.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops

_PERM_CACHE = {}


def shuffle_indices(length, numShuffles, seed=355687):
    """Row permutations the reference draws for kmer-size 1: ``default_rng(355687)`` is re-created
    at every call and ``rng.permutation(x, axis=0)`` is applied numShuffles times in sequence
    (interpretUtils.py:1218-1222), so every query of a given length sees the same permutations."""
    key = (length, numShuffles, seed)
    if key not in _PERM_CACHE:
        rng = np.random.default_rng(seed=seed)
        idx = np.arange(length, dtype=np.int32)
        _PERM_CACHE[key] = np.stack([rng.permutation(idx, axis=0) for _ in range(numShuffles)])
    return _PERM_CACHE[key]


def make_references(x_u8, numShuffles, seed=355687):
    """(Q, L, 4) u8 on device -> (Q*numShuffles, L, 4) shuffled references on device."""
    Q, L, _ = x_u8.shape
    perm = torch.as_tensor(shuffle_indices(L, numShuffles, seed), device=x_u8.device)
    refs = torch.empty((Q * numShuffles, L, 4), dtype=torch.uint8, device=x_u8.device)
    ops.shuffle_rows(x_u8.contiguous(), perm.contiguous(), refs)
    return refs


def pisa_batch(net, x_u8, headID, taskID, numShuffles, refs_u8=None, hypothetical=False):
    """PISA attributions of Q query windows (interpretPisa.py:153-166: metric = leftmost logit).

    Returns (phi (Q, Lin, 4) fp32, input_predictions (Q,), shuffle_predictions (Q, n))."""
    cfg = net.cfg
    Q = x_u8.shape[0]
    x_u8 = x_u8.contiguous()
    if refs_u8 is None:
        refs_u8 = make_references(x_u8, numShuffles)
    col = int(sum(cfg.headTasks[:headID])) + taskID

    def dlogits_fn(wx, wr):
        wr["dlogits"].zero_()
        wr["dlogcounts"].zero_()
        wr["dlogits"][:, 0, col] = 1.0

    dx, wx, wr = net.shap_batch(x_u8, refs_u8, dlogits_fn)
    phi = torch.empty((Q, cfg.inputLength, 4), dtype=torch.float32, device=x_u8.device)
    ops.shap_combine(x_u8, refs_u8, dx, phi, hypothetical)
    vx = wx["logits"][:, 0, col].clone()
    vr = wr["logits"][:, 0, col].reshape(Q, numShuffles).clone()
    return phi, vx, vr
