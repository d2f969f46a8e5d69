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


def make_references(x_u8, numShuffles, seed=355687, kmerSize=1):
    """(Q, L, 4) u8 on device -> (Q*numShuffles, L, 4) shuffled references on device.

    kmerSize 1: the same row permutations for every query (interpretUtils.py:1216-1222, on the
    device).  kmerSize > 1: k-let preserving shuffles of every query, the generator re-seeded per
    query as the reference does (interpretUtils.py:1223-1225) -- host code, like the reference's
    ushuffle, bit-identical to it (``bpb_ushuffle_ohe``)."""
    Q, L, _ = x_u8.shape
    if kmerSize > 1:
        xh = x_u8.cpu().numpy()
        refs = np.concatenate([ops.ushuffle_ohe(xh[q], kmerSize, numShuffles, seed)
                               for q in range(Q)], axis=0)
        return torch.as_tensor(refs, device=x_u8.device)
    perm = torch.as_tensor(shuffle_indices(L, numShuffles, seed), device=x_u8.device)
    refs = torch.empty((Q * numShuffles, L, 4), dtype=torch.uint8, device=x_u8.device)
    ops.shuffle_rows(x_u8.contiguous(), perm.contiguous(), refs)
    return refs


def _receptive_field_net(net):
    """PISA asks for the attribution of ONE output position (the leftmost logit), which only sees
    the first inputLength - outputLength + 1 input bases (the reference stores exactly that slice:
    interpretPisa.py:45, 'receptive field').  The convolutions are 'valid', so running the same
    weights on that prefix with outputLength 1 gives bit-identical logits and multipliers at
    ~2/3 of the work.  The cropped twin shares nothing mutable with ``net``; it is rebuilt when
    the weights changed (``net.step``)."""
    from .engine import NetConfig, SoloNet
    cfg = net.cfg
    twin = getattr(net, "_rf_twin", None)
    if twin is None or twin[0] != net.step:
        rf = cfg.inputLength - cfg.outputLength + 1
        c2 = NetConfig(inputLength=rf, outputLength=1, numFilters=cfg.numFilters,
                       numLayers=cfg.numLayers, inputFilterWidth=cfg.inputFilterWidth,
                       outputFilterWidth=cfg.outputFilterWidth, headTasks=list(cfg.headTasks),
                       precise=cfg.precise)
        t = SoloNet(c2, net.device)
        t.load_state(net.state())
        twin = (net.step, t)
        net._rf_twin = twin
    return twin[1]


def pisa_batch(net, x_u8, headID, taskID, numShuffles, refs_u8=None, hypothetical=False,
               use_graph=True, crop=True, kmerSize=1):
    """PISA attributions of Q query windows (interpretPisa.py:153-166: metric = leftmost logit).

    Returns (phi (Q, Lin, 4) fp32, input_predictions (Q,), shuffle_predictions (Q, n)); phi is
    zero beyond the receptive field of output position 0.

    With ``crop`` the network is evaluated on the receptive field only (exact, see
    ``_receptive_field_net``).  With the standard shuffled references the whole computation --
    reference generation, the two forward passes, the rescale backward and the combiner, ~45
    launches -- is captured once per (Q, numShuffles, head, task) as a CUDA graph and replayed:
    the launches are shorter than the Python launch path, so eager execution is host-bound."""
    cfg = net.cfg
    Q, Lin = x_u8.shape[0], cfg.inputLength
    x_u8 = x_u8.contiguous()
    if refs_u8 is None and kmerSize > 1:
        refs_u8 = make_references(x_u8, numShuffles, kmerSize=kmerSize)
    col = int(sum(cfg.headTasks[:headID])) + taskID
    run = _receptive_field_net(net) if (crop and cfg.outputLength > 1) else net
    rf = run.cfg.inputLength
    dev = x_u8.device

    def dlogits_fn(wx, wr):
        wr["dlogits"].zero_()
        wr["dlogcounts"].zero_()
        wr["dlogits"][:, 0, col] = 1.0

    def compute(x_full, refs_full, phi_full, xc, rc, phic):
        """x_full/refs_full: whole windows; xc/rc/phic: receptive-field buffers (== the full ones
        when not cropping)."""
        if rf != Lin:
            xc.copy_(x_full[:, :rf])
            rc.copy_(refs_full[:, :rf])
        dx, wx, wr = run.shap_batch(xc, rc, dlogits_fn)
        ops.shap_combine(xc, rc, dx, phic, hypothetical)
        if rf != Lin:
            phi_full.zero_()
            phi_full[:, :rf].copy_(phic)
        return wx, wr

    def buffers():
        R = Q * numShuffles
        st = {"x": torch.zeros((Q, Lin, 4), dtype=torch.uint8, device=dev),
              "refs": torch.empty((R, Lin, 4), dtype=torch.uint8, device=dev),
              "phi": torch.empty((Q, Lin, 4), dtype=torch.float32, device=dev)}
        if rf != Lin:
            st["xc"] = torch.empty((Q, rf, 4), dtype=torch.uint8, device=dev)
            st["rc"] = torch.empty((R, rf, 4), dtype=torch.uint8, device=dev)
            st["phic"] = torch.empty((Q, rf, 4), dtype=torch.float32, device=dev)
        else:
            st["xc"], st["rc"], st["phic"] = st["x"], st["refs"], st["phi"]
        return st

    if refs_u8 is not None or not use_graph:
        st = buffers()
        st["x"].copy_(x_u8)
        st["refs"].copy_(refs_u8 if refs_u8 is not None else make_references(x_u8, numShuffles))
        wx, wr = compute(st["x"], st["refs"], st["phi"], st["xc"], st["rc"], st["phic"])
        phi = st["phi"]
    else:
        key = ("pisa", Q, numShuffles, col, bool(hypothetical), rf)
        # cached on the network that actually runs: a rebuilt receptive-field twin (new weights)
        # takes its graphs -- which hold pointers into its buffers -- with it
        st = run._graphs.get(key)
        if st is None:
            st = buffers()
            st["perm"] = torch.as_tensor(shuffle_indices(Lin, numShuffles), device=dev).contiguous()

            def body():
                ops.shuffle_rows(st["x"], st["perm"], st["refs"])
                st["wx"], st["wr"] = compute(st["x"], st["refs"], st["phi"], st["xc"], st["rc"],
                                             st["phic"])

            st["x"].copy_(x_u8)
            body()  # eager warm-up (allocates workspaces, configures kernels)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                body()
            st["graph"] = g
            run._graphs[key] = st
        st["x"].copy_(x_u8)
        st["graph"].replay()
        phi, wx, wr = st["phi"].clone(), st["wx"], st["wr"]
    vx = wx["logits"][:, 0, col].clone()
    vr = wr["logits"][:, 0, col].reshape(Q, numShuffles).clone()
    return phi, vx, vr


def flat_batch(net, x_u8, headID, numShuffles, kind="profile", taskIDs=None, refs_u8=None,
               kmerSize=1):
    """interpretFlat hypothetical contribution scores (interpretFlat.py:187-312,
    interpretUtils.py:1295-1312) of Q windows for one head.

    kind "profile": the weighted mean-normalised logits of ``taskIDs`` (default: every task of the
    head); kind "counts": the head's log-counts.  Returns (hyp_scores (Q, Lin, 4) fp32,
    metric on the queries (Q,), metric on the shuffles (Q, n))."""
    cfg = net.cfg
    Q = x_u8.shape[0]
    x_u8 = x_u8.contiguous()
    dev = x_u8.device
    if refs_u8 is None:
        refs_u8 = make_references(x_u8, numShuffles, kmerSize=kmerSize)
    k0 = int(sum(cfg.headTasks[:headID]))
    tasks = list(range(cfg.headTasks[headID])) if taskIDs is None else list(taskIDs)
    sel = torch.zeros((cfg.T,), dtype=torch.int32, device=dev)
    for t in tasks:
        sel[k0 + t] = 1
    vals = {}

    def dlogits_fn(wx, wr):
        wr["dlogits"].zero_()
        wr["dlogcounts"].zero_()
        if kind == "counts":
            wr["dlogcounts"][:, headID] = 1.0
            vals["x"] = wx["logcounts"][:, headID].clone()
            vals["r"] = wr["logcounts"][:, headID].reshape(Q, numShuffles).clone()
        else:
            vx = torch.empty((Q,), dtype=torch.float32, device=dev)
            vr = torch.empty((Q * numShuffles,), dtype=torch.float32, device=dev)
            ops.flat_profile_seed(sel, wx["logits"], wr["logits"], wr["dlogits"], vx, vr)
            vals["x"], vals["r"] = vx, vr.reshape(Q, numShuffles)

    dx, _, _ = net.shap_batch(x_u8, refs_u8, dlogits_fn)
    hyp = torch.empty((Q, cfg.inputLength, 4), dtype=torch.float32, device=dev)
    ops.shap_combine(x_u8, refs_u8, dx, hyp, True)
    return hyp, vals["x"], vals["r"]
