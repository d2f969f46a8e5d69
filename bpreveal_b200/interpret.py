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


# ---------------------------------------------------------------------------------------------
# transformation / combined models (shap.py:782-793: Exp, Log, Sigmoid under the rescale rule)
# ---------------------------------------------------------------------------------------------
def _model_parts(model):
    """(residual network | None, bias network, transformation tables, crop flank, use-bias flags)
    of a SoloModel / TransformationModel / CombinedModel."""
    from .models import CombinedModel, SoloModel, TransformationModel
    if isinstance(model, SoloModel):
        return None
    if isinstance(model, TransformationModel):
        tm = model
        return (None, tm.solo.net, (tm.pAct, tm.cAct, tm.coeffs, len(tm.pTypes), len(tm.cTypes)),
                0, None)
    if isinstance(model, CombinedModel):
        bnet, pAct, cAct, coeffs, n_p, n_c = model._bias_parts()
        return (model.residual.net, bnet, (pAct, cAct, coeffs, n_p, n_c), model.cropFlank,
                model.useBias)
    raise TypeError(f"cannot interpret {type(model).__name__}")


def model_shap_batch(model, x_u8, refs_u8, seed_fn, hypothetical):
    """DeepSHAP of Q queries against their n references each on ANY of the three model kinds.

    ``seed_fn(logits_x (Q,L,T), lc_x (Q,H), logits_r (R,L,T), lc_r (R,H), dlogits (R,L,T),
    dlc (R,H))`` sees the MODEL's outputs (transformed / combined) on the queries and on the
    references and fills the metric's gradient wrt the outputs of every pair; it returns
    ``(metric on queries (Q,), metric on references (Q, n))``.

    For a combined model the attribution is the sum of two rescale-rule backward passes -- through
    the residual network, and through the frozen bias network on the cropped window (models.py:
    339-347) seeded with the gradient propagated through transformation, add and
    CountsLogSumExp by ``bpb_combined_shap_seed``.  Returns (phi (Q, Lin, 4), vx, vr)."""
    parts = _model_parts(model)
    Q, Lin, _ = x_u8.shape
    R = refs_u8.shape[0]
    n = R // Q
    dev = x_u8.device
    x_u8, refs_u8 = x_u8.contiguous(), refs_u8.contiguous()
    phi = torch.empty((Q, Lin, 4), dtype=torch.float32, device=dev)
    if parts is None:
        net = model.net

        def dlogits_fn(wx, wr):
            wr["dlogits"].zero_()
            wr["dlogcounts"].zero_()
            model_shap_batch.vals = seed_fn(wx["logits"], wx["logcounts"], wr["logits"],
                                            wr["logcounts"], wr["dlogits"], wr["dlogcounts"])
        dx, _, _ = net.shap_batch(x_u8, refs_u8, dlogits_fn)
        ops.shap_combine(x_u8, refs_u8, dx, phi, hypothetical)
        return (phi,) + tuple(model_shap_batch.vals)
    rnet, bnet, (pAct, cAct, coeffs, n_p, n_c), f, useBias = parts
    xb = x_u8[:, f:Lin - f, :].contiguous() if f > 0 else x_u8
    rb = refs_u8[:, f:Lin - f, :].contiguous() if f > 0 else refs_u8
    bx, br = bnet.shap_forward(xb, rb)
    T, H = bnet.cfg.T, bnet.cfg.H
    zeros_l = zeros_c = None
    if rnet is not None:
        rx, rr = rnet.shap_forward(x_u8, refs_u8)
        rl_x, rc_x, rl_r, rc_r = rx["logits"], rx["logcounts"], rr["logits"], rr["logcounts"]
        use = useBias
    else:
        # transformation model alone: "residual" outputs of zero, log-counts not combined
        rl_x = torch.zeros_like(bx["logits"])
        rl_r = torch.zeros_like(br["logits"])
        rc_x = torch.zeros_like(bx["logcounts"])
        rc_r = torch.zeros_like(br["logcounts"])
        use = torch.zeros((H,), dtype=torch.int32, device=dev)
        zeros_l, zeros_c = rl_x, rc_x
    # the model's outputs on queries and references
    ox_l, ox_c = torch.empty_like(bx["logits"]), torch.empty_like(bx["logcounts"])
    or_l, or_c = torch.empty_like(br["logits"]), torch.empty_like(br["logcounts"])
    bt_x, bt_r = torch.empty_like(ox_c), torch.empty_like(or_c)
    ops.combine_bias(bnet.task_off, bx["logits"], bx["logcounts"], pAct, cAct, coeffs, n_p, n_c,
                     use, rl_x, rc_x, ox_l, ox_c, bias_lc_t=bt_x)
    ops.combine_bias(bnet.task_off, br["logits"], br["logcounts"], pAct, cAct, coeffs, n_p, n_c,
                     use, rl_r, rc_r, or_l, or_c, bias_lc_t=bt_r)
    if rnet is None:  # the transformation model's log-counts ARE the transformed bias ones
        ox_c, or_c = bt_x, bt_r
    dl_c = torch.zeros_like(or_l)
    dc_c = torch.zeros_like(or_c)
    vx, vr = seed_fn(ox_l, ox_c, or_l, or_c, dl_c, dc_c)
    ops.combined_shap_seed(n, bnet.task_off, bx["logits"], br["logits"], bx["logcounts"],
                           br["logcounts"], rc_x if rnet is not None else None,
                           rc_r if rnet is not None else None, pAct, cAct, coeffs, n_p, n_c,
                           use, dl_c, dc_c,
                           dl_r=rr["dlogits"] if rnet is not None else None,
                           dc_r=rr["dlogcounts"] if rnet is not None else None,
                           dl_b=br["dlogits"], dc_b=br["dlogcounts"])
    dxb = bnet.shap_backward(bx, br)
    phib = torch.empty((Q, xb.shape[1], 4), dtype=torch.float32, device=dev)
    ops.shap_combine(xb, rb, dxb, phib, hypothetical)
    if rnet is not None:
        dxr = rnet.shap_backward(rx, rr)
        ops.shap_combine(x_u8, refs_u8, dxr, phi, hypothetical)
    else:
        phi.zero_()
    phi[:, f:Lin - f, :] += phib
    del zeros_l, zeros_c
    return phi, vx, vr


def pisa_model(model, x_u8, headID, taskID, numShuffles, refs_u8=None, kmerSize=1):
    """PISA on any model kind (solo models: prefer ``pisa_batch``, which adds the receptive-field
    crop and the CUDA graph).  Returns (phi, input_predictions, shuffle_predictions)."""
    from .models import SoloModel
    if isinstance(model, SoloModel):
        return pisa_batch(model.net, x_u8, headID, taskID, numShuffles, refs_u8=refs_u8,
                          kmerSize=kmerSize)
    Q = x_u8.shape[0]
    if refs_u8 is None:
        refs_u8 = make_references(x_u8.contiguous(), numShuffles, kmerSize=kmerSize)
    col = int(sum(model.headTasks[:headID])) + taskID

    def seed(lx, cx, lr, cr, dl, dc):
        dl[:, 0, col] = 1.0
        return lx[:, 0, col].clone(), lr[:, 0, col].reshape(Q, numShuffles).clone()
    return model_shap_batch(model, x_u8, refs_u8, seed, hypothetical=False)


def flat_model(model, x_u8, headID, numShuffles, kind="profile", taskIDs=None, refs_u8=None,
               kmerSize=1):
    """interpretFlat hypothetical contribution scores on any model kind."""
    from .models import SoloModel
    if isinstance(model, SoloModel):
        return flat_batch(model.net, x_u8, headID, numShuffles, kind=kind, taskIDs=taskIDs,
                          refs_u8=refs_u8, kmerSize=kmerSize)
    Q = x_u8.shape[0]
    dev = x_u8.device
    if refs_u8 is None:
        refs_u8 = make_references(x_u8.contiguous(), numShuffles, kmerSize=kmerSize)
    k0 = int(sum(model.headTasks[:headID]))
    tasks = list(range(model.headTasks[headID])) if taskIDs is None else list(taskIDs)
    T = int(sum(model.headTasks))
    sel = torch.zeros((T,), dtype=torch.int32, device=dev)
    for t in tasks:
        sel[k0 + t] = 1

    def seed(lx, cx, lr, cr, dl, dc):
        if kind == "counts":
            dc[:, headID] = 1.0
            return cx[:, headID].clone(), cr[:, headID].reshape(Q, numShuffles).clone()
        vx = torch.empty((Q,), dtype=torch.float32, device=dev)
        vr = torch.empty((Q * numShuffles,), dtype=torch.float32, device=dev)
        ops.flat_profile_seed(sel, lx.contiguous(), lr.contiguous(), dl, vx, vr)
        return vx, vr.reshape(Q, numShuffles)
    return model_shap_batch(model, x_u8, refs_u8, seed, hypothetical=True)
