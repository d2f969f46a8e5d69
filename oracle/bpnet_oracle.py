"""CPU oracle for the BPReveal BPNet hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU (torch fp64 / fp32 + numpy), the arithmetic that the
reference describes as a Keras graph and executes through TensorFlow.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it; the product package ``bpreveal_b200`` never does.

PARITY UNPINNED for the tensor arithmetic: the reference ships no tests, golden vectors or
fixtures for this path and TensorFlow/Keras are not installable here,
so the conv/loss/SHAP restatement below is pinned only by (i) the reference's own
known-answer material (lengthCalc outputs, oneHotEncode docstring), (ii) the reference's C
helpers compiled from their sources into ``oracle/_ref`` (libslide, libushuffle) and (iii)
mathematical identities (efficiency of DeepSHAP, gradient checks in fp64).

Each function cites the reference file:line (relative to /root/reference/src) it follows.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

NUM_BASES = 4  # internal/constants.py:152


# --------------------------------------------------------------------------------------
# lengthCalc.py:71-123
# --------------------------------------------------------------------------------------
def length_difference(numLayers: int, inputFilterWidth: int, outputFilterWidth: int) -> int:
    """Input length minus output length of a solo model (lengthCalc.py:71-123)."""
    overhang = inputFilterWidth - 1
    for i in range(numLayers):
        # A width-3 filter with dilation 2^(i+1) consumes 2 * 2^(i+1) bases (models.py:93-96).
        overhang += 2 * (2 ** (i + 1))
    overhang += outputFilterWidth - 1
    return overhang


# --------------------------------------------------------------------------------------
# utils.py:420-486  oneHotEncode / oneHotDecode
# --------------------------------------------------------------------------------------
def one_hot_encode(sequence: str, allowN: bool = False) -> np.ndarray:
    """utils.py:420-486.  A->[1,0,0,0], C->[0,1,0,0], G->[0,0,1,0], T->[0,0,0,1], else zeros."""
    out = np.zeros((len(sequence), NUM_BASES), dtype=np.uint8)
    lut = {"A": 0, "C": 1, "G": 2, "T": 3, "a": 0, "c": 1, "g": 2, "t": 3}
    for i, ch in enumerate(sequence):
        j = lut.get(ch)
        if j is not None:
            out[i, j] = 1
        elif not allowN:
            raise AssertionError("Invalid sequence: only ACGT allowed")
    return out


# --------------------------------------------------------------------------------------
# Parameter containers.  Layout follows Keras: conv kernels (k, C_in, C_out), channels-last.
# --------------------------------------------------------------------------------------
def glorot_uniform(rng: np.random.Generator, shape, fan_in, fan_out):
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def init_solo_params(numFilters, numLayers, inputFilterWidth, outputFilterWidth, headTasks,
                     seed=1234, scale=1.0, biasScale=0.0):
    """Random-init parameters of a solo model (Keras defaults: Glorot-uniform kernels, zero bias).

    ``headTasks`` is a list of num-tasks per head.  ``scale`` > 1 gives "trained-like" magnitudes;
    ``biasScale`` > 0 draws non-zero biases so that bias handling is actually exercised in tests.
    """
    rng = np.random.default_rng(seed)
    p = {}
    Fn = numFilters
    p["conv1_w"] = glorot_uniform(rng, (inputFilterWidth, NUM_BASES, Fn),
                                  inputFilterWidth * NUM_BASES, inputFilterWidth * Fn) * scale
    p["conv1_b"] = (rng.standard_normal(Fn) * biasScale).astype(np.float32)
    for i in range(numLayers):
        p[f"dil{i}_w"] = glorot_uniform(rng, (3, Fn, Fn), 3 * Fn, 3 * Fn) * scale
        p[f"dil{i}_b"] = (rng.standard_normal(Fn) * biasScale).astype(np.float32)
    for h, T in enumerate(headTasks):
        p[f"prof{h}_w"] = glorot_uniform(rng, (outputFilterWidth, Fn, T),
                                         outputFilterWidth * Fn, outputFilterWidth * T) * scale
        p[f"prof{h}_b"] = (rng.standard_normal(T) * biasScale).astype(np.float32)
        p[f"cnt{h}_w"] = glorot_uniform(rng, (Fn, 1), Fn, 1) * scale
        p[f"cnt{h}_b"] = (rng.standard_normal(1) * biasScale).astype(np.float32)
    return p


def to_torch(params, dtype=torch.float64, requires_grad=False):
    out = {}
    for k, v in params.items():
        t = torch.tensor(np.asarray(v), dtype=dtype)
        t.requires_grad_(requires_grad)
        out[k] = t
    return out


def _conv1d_cl(x, w, b, dilation=1):
    """Keras Conv1D 'valid', channels-last, cross-correlation.

    x: (B, L, Cin); w: (k, Cin, Cout) -> (B, L - d(k-1), Cout)
    """
    y = F.conv1d(x.transpose(1, 2), w.permute(2, 1, 0), b, dilation=dilation)
    return y.transpose(1, 2)


def num_layers_of(p):
    n = 0
    while f"dil{n}_w" in p:
        n += 1
    return n


def num_heads_of(p):
    n = 0
    while f"prof{n}_w" in p:
        n += 1
    return n


# --------------------------------------------------------------------------------------
# models.py:53-115 soloModel, models.py:20-50 _soloModelHead
# --------------------------------------------------------------------------------------
def solo_forward(x, p, relu=F.relu, keep=False):
    """Forward of soloModel.  x: (B, Lin, 4).  Returns ([profiles...], [logcounts (B,1)...]).

    models.py:87-90 initial conv + relu; models.py:92-104 dilated conv (d = 2^(i+1)) + relu,
    crop of the *previous* layer by (prev-new)//2 = d, add; models.py:39-49 heads.
    ``relu`` is injectable so that the DeepSHAP oracle can swap in the rescale rule.
    With ``keep`` the list of post-add activations and pre-activations is also returned.
    """
    acts, zs = [], []
    z = _conv1d_cl(x, p["conv1_w"], p["conv1_b"])
    a = relu(z)
    zs.append(z)
    acts.append(a)
    for i in range(num_layers_of(p)):
        d = 2 ** (i + 1)
        z = _conv1d_cl(a, p[f"dil{i}_w"], p[f"dil{i}_b"], dilation=d)
        crop = (a.shape[1] - z.shape[1]) // 2
        a = relu(z) + a[:, crop:a.shape[1] - crop, :]
        zs.append(z)
        acts.append(a)
    profiles, counts = [], []
    for h in range(num_heads_of(p)):
        profiles.append(_conv1d_cl(a, p[f"prof{h}_w"], p[f"prof{h}_b"]))
        gap = a.mean(dim=1)  # GlobalAveragePooling1D, models.py:43-45
        counts.append(gap @ p[f"cnt{h}_w"] + p[f"cnt{h}_b"])  # Dense(1), models.py:46-49
    if keep:
        return profiles, counts, acts, zs
    return profiles, counts


def round_bf16(t):
    """Round-to-nearest-even to bfloat16 (8 significand bits), returned in t's dtype."""
    return t.to(torch.float32).to(torch.bfloat16).to(t.dtype)


def solo_forward_bf16_emulation(x, p):
    """The same graph as ``solo_forward`` under the storage rules of a bf16 tensor-core pipeline
    with exact (fp64) accumulation: kernels of the dilated layers and of the heads rounded to
    bf16, every stored activation rounded to bf16, the first layer's kernel kept (its one-hot
    input makes extra weight planes free).  It is NOT a model of any kernel: it is the error
    floor of bf16 storage, which the tests use to judge whether a bf16 kernel path loses more
    than the number format itself does."""
    q = {k: (round_bf16(v) if k.endswith("_w") and not k.startswith("conv1") else v)
         for k, v in p.items()}
    a = round_bf16(F.relu(_conv1d_cl(x, q["conv1_w"], q["conv1_b"])))
    for i in range(num_layers_of(q)):
        d = 2 ** (i + 1)
        z = _conv1d_cl(a, q[f"dil{i}_w"], q[f"dil{i}_b"], dilation=d)
        a = round_bf16(F.relu(z) + a[:, d:a.shape[1] - d, :])
    profiles, counts = [], []
    for h in range(num_heads_of(q)):
        profiles.append(_conv1d_cl(a, q[f"prof{h}_w"], q[f"prof{h}_b"]))
        counts.append(a.mean(dim=1) @ q[f"cnt{h}_w"] + q[f"cnt{h}_b"])
    return profiles, counts


# --------------------------------------------------------------------------------------
# layers.py:10-49 linearRegression, models.py:118-175 simple transformation
# --------------------------------------------------------------------------------------
def init_transform_params(numHeads, profileSpec, countsSpec):
    """Scalars of a transformation model: kernel_initializer Ones, bias zeros (layers.py:40-44)."""
    tp = {}
    for h in range(numHeads):
        for which, spec in (("profile", profileSpec), ("logcounts", countsSpec)):
            if spec["name"] != "simple":
                continue
            for ty in spec["types"]:
                if ty == "linear":
                    tp[f"{which}{h}_linear"] = np.array([1.0, 0.0], dtype=np.float32)
                else:  # sigmoid / relu: in_linear (m2,b2), out_linear (m1,b1)
                    tp[f"{which}{h}_{ty}_in"] = np.array([1.0, 0.0], dtype=np.float32)
                    tp[f"{which}{h}_{ty}_out"] = np.array([1.0, 0.0], dtype=np.float32)
    return tp


def _simple_transform(u, tp, key, spec, relu=F.relu, sigmoid=torch.sigmoid):
    """models.py:118-175: sum over the listed layer types of regressions of u."""
    if spec["name"] == "passthrough":
        return u
    parts = []
    for ty in spec["types"]:
        if ty == "linear":
            m, b = tp[f"{key}_linear"]
            parts.append(u * m + b)
        elif ty in ("sigmoid", "relu"):
            m2, b2 = tp[f"{key}_{ty}_in"]
            m1, b1 = tp[f"{key}_{ty}_out"]
            act = sigmoid if ty == "sigmoid" else relu
            parts.append(act(u * m2 + b2) * m1 + b1)
        else:
            raise ValueError(ty)
    out = parts[0]
    for q in parts[1:]:
        out = out + q
    return out


def transformation_forward(x, solo_p, tp, profileSpec, countsSpec, relu=F.relu,
                           sigmoid=torch.sigmoid):
    """models.py:226-265 transformationModel (solo frozen; outputs [profiles..., counts...])."""
    profs, cnts = solo_forward(x, solo_p, relu=relu)
    po = [_simple_transform(u, tp, f"profile{h}", profileSpec, relu, sigmoid)
          for h, u in enumerate(profs)]
    co = [_simple_transform(u, tp, f"logcounts{h}", countsSpec, relu, sigmoid)
          for h, u in enumerate(cnts)]
    return po, co


# --------------------------------------------------------------------------------------
# models.py:268-387 combinedModel, layers.py:52-75 CountsLogSumExp
# --------------------------------------------------------------------------------------
def counts_logsumexp(a, b, exp=torch.exp, log=torch.log):
    """layers.py:66-75: log(exp(a)+exp(b)) in float64, cast back; deliberately no max trick."""
    dt = a.dtype
    return log(exp(a.double()) + exp(b.double())).to(dt)


def combined_forward(x, resid_p, bias_solo_p, tp, profileSpec, countsSpec, useBiasCounts,
                     biasInputLength, relu=F.relu, sigmoid=torch.sigmoid,
                     exp=torch.exp, log=torch.log):
    """models.py:331-387.  Bias input is x cropped symmetrically (models.py:339-347); profile
    logits are added (363-365); counts are logsumexp'd where use-bias-counts (366-380)."""
    crop = (x.shape[1] - biasInputLength) // 2
    xb = x[:, crop:x.shape[1] - crop, :]
    bprof, bcnt = transformation_forward(xb, bias_solo_p, tp, profileSpec, countsSpec, relu,
                                         sigmoid)
    rprof, rcnt = solo_forward(x, resid_p, relu=relu)
    profs = [bprof[h] + rprof[h] for h in range(len(rprof))]
    cnts = []
    for h in range(len(rcnt)):
        if useBiasCounts[h]:
            cnts.append(counts_logsumexp(bcnt[h], rcnt[h], exp, log))
        else:
            cnts.append(rcnt[h])
    return profs, cnts, (rprof, rcnt)


# --------------------------------------------------------------------------------------
# losses.py:15-39 multinomialNll ; losses.py:42-63 weightedMse (+ Keras reduction)
# --------------------------------------------------------------------------------------
def multinomial_nll(trueCounts, logits):
    """losses.py:15-39 with tfp Multinomial.log_prob:
    log_prob = lgamma(N+1) - sum lgamma(k+1) + sum k * log_softmax(logits), over L*T flattened.
    Loss = -sum_b log_prob / B.
    """
    B = trueCounts.shape[0]
    k = trueCounts.reshape(B, -1)
    lg = logits.reshape(B, -1)
    N = k.sum(dim=1)
    logp = torch.lgamma(N + 1) - torch.lgamma(k + 1).sum(dim=1) \
        + (k * torch.log_softmax(lg, dim=1)).sum(dim=1)
    return -logp.sum() / B


def weighted_mse(yTrue, yPred, lam):
    """losses.py:57-62 + Keras: yTrue (B,) is expanded to (B,1) to match yPred (B,1)
    (squeeze_or_expand_to_same_rank), mean over the last axis, times lambda, then
    sum_over_batch_size reduction = mean over the batch."""
    if yTrue.dim() == yPred.dim() - 1:
        yTrue = yTrue.unsqueeze(-1)
    mse = ((yTrue - yPred) ** 2).mean(dim=-1)
    return (mse * lam).mean()


def total_loss(profs, cnts, trueProfiles, trueLogCounts, profileWeights, lambdas):
    """training.py:64-65 + trainSoloModel.py:183-185: sum_h w_h * mnll_h + sum_h 1.0 * wmse_h.
    Returns (total, [mnll...], [wmse...])."""
    pl = [multinomial_nll(trueProfiles[h], profs[h]) for h in range(len(profs))]
    cl = [weighted_mse(trueLogCounts[h], cnts[h], lambdas[h]) for h in range(len(cnts))]
    tot = sum(w * l for w, l in zip(profileWeights, pl)) + sum(cl)
    return tot, pl, cl


# --------------------------------------------------------------------------------------
# Keras-3 Adam (recalled semantics; not in /root/reference)
# --------------------------------------------------------------------------------------
def adam_step(param, grad, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-7):
    """One Keras Adam update in place on numpy/torch arrays.  ``step`` is 1-based.

    alpha = lr * sqrt(1 - b2^t) / (1 - b1^t); m += (g - m)(1 - b1); v += (g^2 - v)(1 - b2);
    theta -= alpha * m / (sqrt(v) + eps)   (epsilon outside the sqrt, not bias corrected).
    """
    alpha = lr * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
    m += (grad - m) * (1.0 - beta1)
    v += (grad * grad - v) * (1.0 - beta2)
    param -= alpha * m / (v ** 0.5 + eps)
    return param, m, v


# --------------------------------------------------------------------------------------
# generators.py:108-140 + libslide.c:44-91
# --------------------------------------------------------------------------------------
def slide(source, dest, rowIndexes, colIndexes):
    """libslide.c:44-66: dest[rowIndexes[r]] = source[r, colIndexes[r]:colIndexes[r]+numDestCols]."""
    n = dest.shape[1]
    for r in range(source.shape[0]):
        dest[rowIndexes[r]] = source[r, colIndexes[r]:colIndexes[r] + n]
    return dest


def epoch_jitter(numRegions, maxJitter, rng):
    """generators.py:108-122: shuffle region order in place, draw offsets in [0, 2*maxJitter)."""
    regionIndexes = np.arange(0, numRegions)
    rng.shuffle(regionIndexes)
    sliceCols = rng.integers(0, maxJitter * 2, size=(numRegions,), dtype=np.int32)
    return regionIndexes, sliceCols


def log_counts_targets(values):
    """generators.py:139-140: log(sum over (length, tasks))."""
    return np.log(values.sum(axis=(1, 2)))


# --------------------------------------------------------------------------------------
# interpretUtils.py:1216-1225 generateShuffles (kmer 1)
# --------------------------------------------------------------------------------------
def generate_shuffles(x, numShuffles, seed=355687):
    """The reference re-creates default_rng(355687) at every call and permutes rows
    ``numShuffles`` times sequentially (interpretUtils.py:1218-1222)."""
    rng = np.random.default_rng(seed=seed)
    return np.array([rng.permutation(x, axis=0) for _ in range(numShuffles)])


def shuffle_indices(length, numShuffles, seed=355687):
    """Same permutations as ``generate_shuffles`` expressed as row indices (identical for every
    query of a given length because the generator is re-seeded per call)."""
    rng = np.random.default_rng(seed=seed)
    idx = np.arange(length)
    return np.array([rng.permutation(idx, axis=0) for _ in range(numShuffles)])


# --------------------------------------------------------------------------------------
# shap.py:612-639 nonlinearity_1d_handler (rescale rule) as a torch autograd function
# --------------------------------------------------------------------------------------
def _make_rescale(fn, dfn):
    class _Rescale(torch.autograd.Function):
        """Joint batch [x-half | r-half]; backward replaces the local derivative of both halves
        by (out_x - out_r)/(in_x - in_r) where |in_x - in_r| >= 1e-6, else the ordinary
        derivative of that half (shap.py:623-638)."""

        @staticmethod
        def forward(ctx, z):
            out = fn(z)
            ctx.save_for_backward(z, out)
            return out

        @staticmethod
        def backward(ctx, g):
            z, out = ctx.saved_tensors
            n = z.shape[0] // 2
            din = z[:n] - z[n:]
            dout = out[:n] - out[n:]
            safe = torch.where(din.abs() < 1e-6, torch.ones_like(din), din)
            mult = (dout / safe).repeat((2,) + (1,) * (z.dim() - 1))
            small = (din.abs() < 1e-6).repeat((2,) + (1,) * (z.dim() - 1))
            return torch.where(small, g * dfn(z, out), g * mult)

    return _Rescale.apply


rescale_relu = _make_rescale(F.relu, lambda z, o: (z > 0).to(z.dtype))
rescale_sigmoid = _make_rescale(torch.sigmoid, lambda z, o: o * (1 - o))
rescale_exp = _make_rescale(torch.exp, lambda z, o: o)
rescale_log = _make_rescale(torch.log, lambda z, o: 1.0 / z)


# --------------------------------------------------------------------------------------
# Metrics: interpretPisa.py:153-166 ; interpretFlat.py:187-250
# --------------------------------------------------------------------------------------
def pisa_metric(profs, cnts, headID, taskID):
    """interpretPisa.py:166: model.outputs[head][:, 0, task] (leftmost output logit)."""
    return profs[headID][:, 0, taskID]


def counts_metric(profs, cnts, headID):
    """interpretFlat.py:240-250: model.outputs[numHeads + head][:, 0]."""
    return cnts[headID][:, 0]


class _WeightedSumMul(torch.autograd.Function):
    """interpretFlat.py:233 ``softmaxOut * meannormedLogits`` under shap.py:642-678
    (nonlinearity_2d_handler).  Both inputs vary; the softmax branch is behind a StopGradient so
    only the gradient wrt the logits input is propagated.  For op_func = mul the rule gives
    out1 = 0.5 (p_x l_x - p_x l_r + p_r l_x - p_r l_r) / (l_x - l_r) = 0.5 (p_x + p_r),
    zeroed where |l_x - l_r| < 1e-7."""

    @staticmethod
    def forward(ctx, p, l):
        ctx.save_for_backward(p, l)
        return p * l

    @staticmethod
    def backward(ctx, g):
        p, l = ctx.saved_tensors
        n = p.shape[0] // 2
        dl = l[:n] - l[n:]
        m = 0.5 * (p[:n] + p[n:])
        m = torch.where(dl.abs() < 1e-7, torch.zeros_like(m), m)
        rep = (2,) + (1,) * (p.dim() - 1)
        return None, g * m.repeat(rep)


def profile_metric(profs, cnts, headID, taskIDs, shap=False):
    """interpretFlat.py:187-237: logits of the selected tasks, mean-normalised over the whole
    (length x tasks) block, weighted by the stop-gradient softmax over the same block, summed."""
    lg = profs[headID][:, :, taskIDs]
    B = lg.shape[0]
    flat = lg.reshape(B, -1)
    mn = flat - flat.mean(dim=1, keepdim=True)
    sm = torch.softmax(mn.detach(), dim=1)
    if shap:
        prod = _WeightedSumMul.apply(sm, mn)
    else:
        prod = sm * mn
    return prod.sum(dim=1)


# --------------------------------------------------------------------------------------
# shap.py:347-394 shap_values + shap.py:26-30 / interpretUtils.py:1295-1312 combiners
# --------------------------------------------------------------------------------------
def deepshap(x_onehot, refs, forward_fn, metric_fn, dtype=torch.float64, hypothetical=False):
    """DeepSHAP (rescale) attribution of one sequence against ``refs``.

    x_onehot: (L,4) array; refs: (n,L,4).  ``forward_fn(x, relu=..., ...)`` must build the graph
    with the injected nonlinearities; ``metric_fn(profs, cnts)`` -> (2n,) scalar per sample.
    Returns (phi (L,4), metric on x, metric on refs (n,)).
    shap.py:362-377: joint = [tile(x, n); refs]; grads of the x half; combiner.
    """
    n = refs.shape[0]
    xt = torch.tensor(np.asarray(x_onehot), dtype=dtype)
    rt = torch.tensor(np.asarray(refs), dtype=dtype)
    joint = torch.cat([xt.unsqueeze(0).repeat(n, 1, 1), rt], dim=0).requires_grad_(True)
    out = forward_fn(joint)
    val = metric_fn(*out[:2])
    (g,) = torch.autograd.grad(val.sum(), joint)
    mult = g[:n]
    if not hypothetical:
        phi = (mult * (xt.unsqueeze(0) - rt)).mean(dim=0)  # shap.py:26-30
    else:
        # interpretUtils.py:1295-1312
        proj = torch.zeros_like(rt)
        for i in range(NUM_BASES):
            hyp = torch.zeros_like(xt)
            hyp[:, i] = 1.0
            proj[:, :, i] = ((hyp.unsqueeze(0) - rt) * mult).sum(dim=-1)
        phi = proj.mean(dim=0)
    return phi.detach().numpy(), val[0].item(), val[n:].detach().numpy()


def solo_shap_forward(p):
    """Forward closure of a solo model with the rescale rule at every ReLU."""
    def fwd(joint):
        return solo_forward(joint, p, relu=rescale_relu)
    return fwd


def combined_shap_forward(resid_p, bias_solo_p, tp, profileSpec, countsSpec, useBiasCounts,
                          biasInputLength):
    """Combined graph with rescale at Relu/Sigmoid/Exp/Log (shap.py:782-793)."""
    def fwd(joint):
        return combined_forward(joint, resid_p, bias_solo_p, tp, profileSpec, countsSpec,
                                useBiasCounts, biasInputLength, relu=rescale_relu,
                                sigmoid=rescale_sigmoid, exp=rescale_exp, log=rescale_log)
    return fwd


# --------------------------------------------------------------------------------------
# utils.py:523-572 logitsToProfile
# --------------------------------------------------------------------------------------
def logits_to_profile(logits, logCounts):
    """utils.py:523-572: softmax over the whole (length x tasks) block times exp(logcounts)."""
    lg = np.asarray(logits, dtype=np.float64)
    e = np.exp(lg - lg.max())
    return (e / e.sum()) * np.exp(logCounts)


# --------------------------------------------------------------------------------------
# Synthetic data
# --------------------------------------------------------------------------------------
def synthetic_sequences(n, length, seed=20261019):
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, 4, size=(n, length))
    out = np.zeros((n, length, 4), dtype=np.uint8)
    np.put_along_axis(out, idx[..., None], 1, axis=2)
    return out


def synthetic_profiles(n, length, tasks, seed=20261019, meanReads=200.0):
    """Poisson counts around a smooth random bump mixture; guaranteed sum > 0."""
    rng = np.random.default_rng(seed + 1)
    pos = np.arange(length)[None, :, None]
    centers = rng.uniform(0, length, size=(n, 1, tasks))
    widths = rng.uniform(length / 40, length / 6, size=(n, 1, tasks))
    rate = np.exp(-0.5 * ((pos - centers) / widths) ** 2) + 0.05
    rate = rate / rate.sum(axis=1, keepdims=True) * meanReads
    out = rng.poisson(rate).astype(np.float32)
    out[:, 0, :] += (out.sum(axis=(1, 2)) == 0)[:, None]
    return out
