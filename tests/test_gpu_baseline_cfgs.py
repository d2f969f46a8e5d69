"""GPU parity at the geometries BASELINE.json names (C1 .. C5), against the fp64 CPU oracle.

What the small nets of test_gpu_model.py never reach is covered here: nine / eleven dilated layers
(dilations up to 512 / 2048: the SLAB-staged branch of the convolution kernel with its residual
ring), F = 512 (256-column accumulator tiles, streamed weights), four heads of two tasks (a packed
head gradient of 16 columns), the combined model's backward through the fp64 log-sum-exp, and
PISA with 20 shuffled references at 3090 bp.

Tolerances.  The fp32-accumulate ("precise") path is held to north_star's figures: logits /
log-counts 1e-4 relative, losses 1e-3.  The bf16 path is held to the error floor of bf16 storage
itself (``bf16_floor``: an ideal bf16 pipeline emulated by the oracle loses 1.5e-2 of the largest
logit over C1's eleven layers, more than north_star's 1e-2, so no bf16 kernel can meet that figure
at this depth): measured error <= 1.5 x floor, plus absolute ceilings.  Gradient and attribution
tolerances are stated where they are asserted.  Every measured error is also appended to
``gpurun_out/parity_r2.jsonl`` (when that directory exists) so that the achieved numbers, not
only pass / fail, are on record (profiles/parity_r2.json is a committed copy).
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import bpnet_oracle as O
from tests import helpers as Hh

pytestmark = pytest.mark.gpu

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def record(case, **vals):
    d = os.path.join(_ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_r2.jsonl"), "a", encoding="utf-8") as fp:
            fp.write(json.dumps({"case": case, **{k: float(v) for k, v in vals.items()}}) + "\n")


def cfg_of(numFilters, numLayers, wi, wo, lout, headTasks):
    return dict(inputLength=lout + O.length_difference(numLayers, wi, wo), outputLength=lout,
                numFilters=numFilters, numLayers=numLayers, inputFilterWidth=wi,
                outputFilterWidth=wo, headTasks=list(headTasks))


C1 = cfg_of(64, 9, 25, 23, 1000, [2])            # configs[0]: 3090 -> 1000
C1W75 = cfg_of(64, 9, 25, 75, 1000, [2])         # north_star's 25/75 variant: 3142 -> 1000
C3R = cfg_of(64, 9, 7, 7, 1000, [2, 2, 2, 2])    # configs[2] residual: 3056 -> 1000, 4 heads x 2
C4S = cfg_of(512, 11, 25, 75, 64, [2])           # configs[3] truncated to 64 outputs (8350 in)


def build(cfg, precise, dev, seed=1234, scale=1.0):
    from bpreveal_b200.engine import NetConfig, SoloNet
    net = SoloNet(NetConfig(precise=precise, **cfg), dev)
    params = Hh.oracle_params(cfg, seed=seed, scale=scale)
    net.load_state(params)
    return net, params


def grad_errors(got, ref):
    gscale = max(np.abs(v).max() for v in ref.values())
    return {k: float(np.abs(got[k] - ref[k]).max() / max(np.abs(ref[k]).max(), 1e-3 * gscale))
            for k in ref}


def bf16_floor(params, x, counts, pw, lam):
    """Errors of an IDEAL bf16 pipeline (oracle.solo_forward_bf16_emulation: bf16 weights and
    stored activations, exact accumulation) against the fp64 oracle, in the measures used below.
    Eleven residual layers deep, bf16's 8 significand bits alone cost ~1.5e-2 of the largest
    logit at Keras' initialisation -- above north_star's 1e-2 -- so the bf16 path is gated on
    "no worse than 1.5x what the number format costs" and on absolute ceilings, and the
    north_star figures are asserted on the fp32-accumulate path."""
    p64 = O.to_torch(params, torch.float64)
    xt = torch.tensor(np.asarray(x), dtype=torch.float64)
    ref_p, ref_c = O.solo_forward(xt, p64)
    emu_p, emu_c = O.solo_forward_bf16_emulation(xt, p64)
    ct = torch.tensor(np.asarray(counts), dtype=torch.float64)
    trueP, trueC, k = [], [], 0
    for pr in ref_p:
        T = pr.shape[2]
        trueP.append(ct[:, :, k:k + T])
        trueC.append(torch.log(ct[:, :, k:k + T].sum(dim=(1, 2))))
        k += T
    _, pl0, cl0 = O.total_loss(ref_p, ref_c, trueP, trueC, pw, lam)
    _, pl1, cl1 = O.total_loss(emu_p, emu_c, trueP, trueC, pw, lam)
    return dict(
        logits=Hh.rel_err(torch.cat(emu_p, 2).numpy(), torch.cat(ref_p, 2).numpy()),
        logcounts=Hh.rel_err(torch.cat(emu_c, 1).numpy(), torch.cat(ref_c, 1).numpy()),
        mnll=max(abs(a.item() - b.item()) / abs(b.item()) for a, b in zip(pl1, pl0)),
        wmse=max(abs(a.item() - b.item()) / abs(b.item()) for a, b in zip(cl1, cl0)))


def gate_bf16(name, got, floor, ceil):
    """got <= max(1.5 x the bf16 floor, 2e-3) and got <= the absolute ceiling."""
    for k, c in ceil.items():
        lim = min(max(1.5 * floor[k], 2e-3), c)
        assert got[k] <= lim, f"{name} {k}: rel err {got[k]:.3e} > {lim:.3e} " \
                              f"(bf16 floor {floor[k]:.3e}, ceiling {c:.1e})"


def test_c1_shape_is_the_baseline_shape():
    assert C1["inputLength"] == 3090 and C1W75["inputLength"] == 3142
    assert C3R["inputLength"] == 3056 and C4S["inputLength"] == 8350
    assert cfg_of(512, 11, 25, 75, 1000, [2])["inputLength"] == 9286


# ------------------------------------------------------------------------------------------
# (i) C1 at full size: forward, losses, every gradient, three Adam steps
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,cfg", [("C1", C1), ("C1w75", C1W75)])
@pytest.mark.parametrize("precise", [False, True])
def test_c1_forward_loss_gradients(dev, name, cfg, precise):
    if name == "C1w75" and precise:
        pytest.skip("the 25/75 variant is checked on the headline (bf16) path only")
    # Keras' own initialisation (Glorot-uniform, scale 1): the state the training benchmark runs
    # on.  Logits are O(1..7) there; "trained-like" x1.5 weights compound over eleven layers to
    # logits of +-80, a regime no trained BPNet is in.
    net, params = build(cfg, precise, dev, scale=1.0)
    net.enable_training()
    B = 3
    x = O.synthetic_sequences(B, cfg["inputLength"], seed=41)
    counts = O.synthetic_profiles(B, 1000, 2, seed=41, meanReads=200.0)
    pw, lam = [1.0], [7.3]
    net.profile_weight.copy_(torch.tensor(pw))
    net.lam.copy_(torch.tensor(lam))
    ref_l, ref_c = Hh.oracle_forward(params, x)
    tot, pl, cl, grads = Hh.oracle_loss_and_grads(params, x, counts, pw, lam)
    got_l, got_c = net.predict(torch.tensor(x, device=dev), batchSize=2)
    el, ec = Hh.rel_err(got_l.cpu().numpy(), ref_l), Hh.rel_err(got_c.cpu().numpy(), ref_c)
    ws = net._workspace(B, "train")
    ws["x"].copy_(torch.tensor(x, device=dev))
    ws["counts"].copy_(torch.tensor(counts, device=dev))
    net.forward_backward_into(ws)
    torch.cuda.synchronize()
    losses = ws["losses"].cpu().numpy()
    e_mnll = abs(losses[0] - pl[0]) / abs(pl[0])
    e_mse = abs(losses[1] - cl[0]) / abs(cl[0])
    ge = grad_errors(net.state(net.grads_flat), grads)
    got = dict(logits=el, logcounts=ec, mnll=e_mnll, wmse=e_mse)
    floor = {} if precise else bf16_floor(params, x, counts, pw, lam)
    record(f"{name}_fwd_bwd_{'precise' if precise else 'bf16'}", grad_max=max(ge.values()),
           grad_dil8=ge["dil8_w"], grad_conv1=ge["conv1_w"], **got,
           **{f"floor_{k}": v for k, v in floor.items()})
    if precise:  # north_star: 1e-4 logits / log-counts, 1e-3 losses
        assert el < 1e-4, f"logits rel err {el}"
        assert ec < 1e-4, f"logcounts rel err {ec}"
        assert e_mnll < 1e-3, f"mnll rel err {e_mnll}"
        assert e_mse < 1e-3, f"weighted mse rel err {e_mse}"
    else:
        gate_bf16(name, got, floor, dict(logits=3e-2, logcounts=5e-2, mnll=1e-2, wmse=2e-2))
    # gradients: relative to each tensor's largest entry.  bf16 path: activations AND loss
    # gradients are rounded to 8 mantissa bits at every layer of an 11-deep chain.
    gtol = 2e-3 if precise else 5e-2
    for k, e in ge.items():
        assert e < gtol, f"grad {k}: rel err {e}"


@pytest.mark.parametrize("precise", [False, True])
def test_c1_three_adam_steps_follow_the_oracle(dev, precise):
    cfg = C1
    net, params = build(cfg, precise, dev)
    B, lr = 2, 0.004
    p = O.to_torch(params, torch.float64, requires_grad=True)
    m = {k: torch.zeros_like(v) for k, v in p.items()}
    v_ = {k: torch.zeros_like(v) for k, v in p.items()}
    worst, first = 0.0, None
    for step in range(1, 4):
        x = O.synthetic_sequences(B, cfg["inputLength"], seed=200 + step)
        counts = O.synthetic_profiles(B, 1000, 2, seed=200 + step, meanReads=150.0)
        ct = torch.tensor(counts, dtype=torch.float64)
        profs, cnts = O.solo_forward(torch.tensor(x, dtype=torch.float64), p)
        tot, pl, cl = O.total_loss(profs, cnts, [ct], [torch.log(ct.sum(dim=(1, 2)))], [1.0],
                                   [1.0])
        for t in p.values():
            t.grad = None
        tot.backward()
        with torch.no_grad():
            for k in p:
                O.adam_step(p[k], p[k].grad, m[k], v_[k], step, lr)
        losses = net.train_step(torch.tensor(x, device=dev), torch.tensor(counts, device=dev), lr)
        got = losses.cpu().numpy()
        e = max(abs(got[0] - pl[0].item()) / abs(pl[0].item()),
                abs(got[1] - cl[0].item()) / abs(cl[0].item()))
        worst = max(worst, e)
        first = e if first is None else first
    got = net.state()
    # Adam normalises every step to ~lr whatever the gradient's size, so a parameter whose
    # gradient is tiny and noisy can move by lr in either direction: compare against lr * steps
    perr = max(float(np.abs(got[k] - p[k].detach().numpy()).max()) for k in params)
    record(f"C1_adam3_{'precise' if precise else 'bf16'}", loss_step1=first, loss_curve=worst,
           param_abs=perr)
    # step 1 is a plain forward pass: the north_star loss tolerance applies on the fp32-accumulate
    # path (bf16: the storage floor of ``bf16_floor``, ~5e-3 here).  From step 2 on the
    # comparison is ill-conditioned by construction: Adam's first updates are lr * sign(g) for
    # EVERY parameter, so each near-zero gradient entry whose sign differs in the last bits moves
    # its weight by 2 lr relative to the oracle (120 k parameters, lr = 0.004 against weights of
    # ~0.1) -- even an fp32 run leaves the fp64 curve by ~1e-3 at step 3.
    assert first < (1e-3 if precise else 1e-2), f"first-step loss rel err {first}"
    assert worst < (3e-3 if precise else 6e-2), f"loss curve rel err {worst}"
    assert perr <= 1.05 * 2 * 3 * lr, f"parameter drift {perr} beyond 2 lr per step"


# ------------------------------------------------------------------------------------------
# (ii) single layers at the dilations / widths the small nets never reach (was: diag_kernels.py)
# ------------------------------------------------------------------------------------------
def split(v, precise):
    hi = v.to(torch.bfloat16)
    lo = (v - hi.float()).to(torch.bfloat16) if precise else None
    return (hi.contiguous(), lo.contiguous() if lo is not None else None)


def val(ps):
    from bpreveal_b200 import ops
    return ops.planes_to_float(ps).double().cpu()


def w_eff(w, precise):
    hi = w.to(torch.bfloat16).float()
    return (hi + (w - hi).to(torch.bfloat16).float() if precise else hi).double().cpu()


def prep_w(w, Fp, precise, dev):
    from bpreveal_b200 import ops
    n = 2 if precise else 1
    fwd = torch.empty((n, 3, Fp, Fp), dtype=torch.bfloat16, device=dev)
    bwd = torch.empty((n, 3, Fp, Fp), dtype=torch.bfloat16, device=dev)
    ops.prep_conv3_weights(w.contiguous(), Fp, fwd, bwd, precise)
    return fwd, bwd


def pack_mask(maskb):
    B, L, Fn = maskb.shape
    mask = torch.zeros((B, L, Fn // 64), dtype=torch.int64, device=maskb.device)
    for c in range(Fn):
        mask[..., c // 64] |= maskb[..., c].long() << (c % 64)
    return mask


LAYER_CASES = [  # (B, L_in, F, dilation, precise)
    (2, 700, 64, 64, False), (2, 900, 64, 128, False), (2, 1300, 64, 256, False),
    (3, 2046, 64, 512, False), (2, 900, 64, 128, True), (2, 2046, 64, 512, True),
    (1, 600, 512, 2, False), (1, 700, 512, 64, False), (1, 900, 512, 128, False),
    (1, 1400, 512, 512, False), (1, 4400, 512, 2048, False), (1, 700, 512, 128, True),
    (2, 700, 256, 128, False), (2, 600, 192, 64, False),
]


def _ids(cases):
    return ["-".join(str(int(v)) if not isinstance(v, bool) else ("p" if v else "b")
                     for v in c) for c in cases]


@pytest.mark.parametrize("B,L,Fn,d,precise", LAYER_CASES, ids=_ids(LAYER_CASES))
def test_dilated_layer_forward(dev, B, L, Fn, d, precise):
    """z = a (*)_d W + b, out = relu(z) + a[d:-d], 1-bit mask (models.py:92-104)."""
    from bpreveal_b200 import ops
    g = torch.Generator().manual_seed(1)
    x = torch.randn((B, L, Fn), generator=g).to(dev)
    w = (torch.randn((3, Fn, Fn), generator=g) / (3 * Fn) ** 0.5).to(dev)
    bias = (torch.randn((Fn,), generator=g) * 0.1).to(dev)
    xin = split(x, precise)
    fwd, _ = prep_w(w, Fn, precise, dev)
    L_out = L - 2 * d
    out = ops.new_planes(B, L_out, Fn, precise, dev)
    mask = torch.zeros((B, L_out, Fn // 64), dtype=torch.int64, device=dev)
    z = torch.empty((B, L_out, Fn), dtype=torch.float32, device=dev)
    ops.conv3(xin, fwd, (0, d, 2 * d), 0, L_out, bias=bias, resid=xin, resid_off=d, out=out,
              mask_out=mask, z_out=z)
    torch.cuda.synchronize()
    xe = val(xin)
    zr = O._conv1d_cl(xe, w_eff(w, precise), bias.double().cpu(), dilation=d)
    ref = torch.relu(zr) + xe[:, d:L - d, :]
    ez = Hh.rel_err(z.cpu().numpy(), zr.numpy())
    eo = Hh.rel_err(val(out).numpy(), ref.numpy())
    assert ez < 1e-4, f"pre-activation rel err {ez}"          # fp32 accumulation in TMEM
    assert eo < (2e-5 if precise else 6e-3), f"output rel err {eo}"  # bf16 rounding of the store
    bits = torch.stack([(mask[..., c // 64] >> (c % 64)) & 1 for c in range(Fn)], dim=-1)
    assert int((bits.bool() != (z > 0)).sum()) == 0, "mask disagrees with the kernel's own z"


@pytest.mark.parametrize("B,L,Fn,d,precise", LAYER_CASES, ids=_ids(LAYER_CASES))
@pytest.mark.parametrize("rule", ["mask", "mult"])
def test_dilated_layer_data_gradient(dev, B, L, Fn, d, precise, rule):
    """g_in = g'[. - d] + sum_j gz[. - j d] W_j^T, gz_in = g_in * relu'(z) (training) or
    g_in * multiplier (DeepSHAP rescale rule, shap.py:612-639)."""
    from bpreveal_b200 import ops
    if rule == "mult" and Fn == 512 and d < 2048 and d != 128:
        pytest.skip("the multiplier epilogue at F = 512 is covered at d = 128 and 2048")
    g = torch.Generator().manual_seed(2)
    L_g = L - 2 * d
    gz = torch.randn((B, L_g, Fn), generator=g).to(dev)
    gp = torch.randn((B, L_g, Fn), generator=g).to(dev)
    w = (torch.randn((3, Fn, Fn), generator=g) / (3 * Fn) ** 0.5).to(dev)
    gzp, gpp = split(gz, precise), split(gp, precise)
    _, bwd = prep_w(w, Fn, precise, dev)
    out = ops.new_planes(B, L, Fn, precise, dev)
    out2 = ops.new_planes(B, L, Fn, precise, dev)
    maskb = (torch.rand((B, L, Fn), generator=g) > 0.5).to(dev)
    mask = pack_mask(maskb)
    mult = split(torch.rand((B, L, Fn), generator=g).to(dev), precise)
    ops.conv3(gzp, bwd, (0, -d, -2 * d), 1, L, resid=gpp, resid_off=-d, out=out, out2=out2,
              mask_in=mask if rule == "mask" else None, mult=mult if rule == "mult" else None)
    torch.cuda.synchronize()
    a = torch.zeros((B, L, Fn), dtype=torch.float64, requires_grad=True)
    zf = O._conv1d_cl(a, w_eff(w, precise), None, dilation=d)
    (zf * val(gzp)).sum().backward()
    ref = a.grad.clone()
    ref[:, d:L - d, :] += val(gpp)
    ref2 = ref * (val(mult) if rule == "mult" else maskb.double().cpu())
    tol = 2e-5 if precise else 6e-3
    e1, e2 = Hh.rel_err(val(out).numpy(), ref.numpy()), Hh.rel_err(val(out2).numpy(), ref2.numpy())
    assert e1 < tol, f"g rel err {e1}"
    assert e2 < tol, f"gz rel err {e2}"


@pytest.mark.parametrize("B,L,Fn,d,precise", LAYER_CASES, ids=_ids(LAYER_CASES))
def test_dilated_layer_weight_gradient(dev, B, L, Fn, d, precise):
    """dW_j = sum_{b,t} a[b, t + j d, :]^T gz[b, t, :], db = sum gz."""
    from bpreveal_b200 import ops
    g = torch.Generator().manual_seed(3)
    L_out = L - 2 * d
    a = split(torch.randn((B, L, Fn), generator=g).to(dev), precise)
    gz = split(torch.randn((B, L_out, Fn), generator=g).to(dev), precise)
    dw = torch.empty((3, Fn, Fn), dtype=torch.float32, device=dev)
    db = torch.empty((Fn,), dtype=torch.float32, device=dev)
    ws = torch.empty((ops.wgrad3_ws_bytes(B, L_out, Fn, precise) // 4,), dtype=torch.float32,
                     device=dev)
    ops.wgrad3(a, gz, d, dw, db, ws)
    torch.cuda.synchronize()
    av, gv = val(a), val(gz)
    ref = torch.stack([torch.einsum("btc,btn->cn", av[:, j * d:j * d + L_out, :], gv)
                       for j in range(3)])
    tol = 3e-5 if precise else 2e-3
    assert Hh.rel_err(dw.cpu().numpy(), ref.numpy()) < tol
    assert Hh.rel_err(db.cpu().numpy(), gv.sum(dim=(0, 1)).numpy()) < tol


XF_CASES = [  # (B, L_in, dilation): F = 64, bf16 -- HALO boxes (d <= 64) and SLAB tiles (d >= 128)
    (2, 300, 2), (2, 421, 8), (2, 700, 64), (2, 900, 128), (2, 1301, 256), (4, 2047, 512),
    (64, 3066, 2), (64, 2046, 512),
]


@pytest.mark.parametrize("B,L,d", XF_CASES, ids=["-".join(map(str, c)) for c in XF_CASES])
def test_dilated_layer_data_gradient_masked_operand(dev, B, L, d):
    """The training backward without gz planes: the operand of the data gradient is the UNMASKED
    gradient g' of the layer above, ANDed with that layer's ReLU bits in shared memory
    (bpb_conv3_args.mask_a):  g_in = g'[. - d] + sum_j (g' (.) bits)[. - j d] W_j^T.
    Odd row counts exercise the 16-byte alignment of the mask-word bulk copies."""
    from bpreveal_b200 import ops
    Fn = 64
    g = torch.Generator().manual_seed(12)
    L_g = L - 2 * d
    gp = torch.randn((B, L_g, Fn), generator=g).to(dev)
    w = (torch.randn((3, Fn, Fn), generator=g) / (3 * Fn) ** 0.5).to(dev)
    gpp = split(gp, False)
    _, bwd = prep_w(w, Fn, False, dev)
    out = ops.new_planes(B, L, Fn, False, dev)
    bits = (torch.rand((B, L_g, Fn), generator=g) > 0.4).to(dev)
    ops.conv3(gpp, bwd, (0, -d, -2 * d), 1, L, resid=gpp, resid_off=-d, out=out,
              mask_a=pack_mask(bits))
    torch.cuda.synchronize()
    a = torch.zeros((B, L, Fn), dtype=torch.float64, requires_grad=True)
    zf = O._conv1d_cl(a, w_eff(w, False), None, dilation=d)
    (zf * (val(gpp) * bits.double().cpu())).sum().backward()
    ref = a.grad.clone()
    ref[:, d:L - d, :] += val(gpp)
    e1 = Hh.rel_err(val(out).numpy(), ref.numpy())
    assert e1 < 6e-3, f"g rel err {e1}"
    # the materialised path on gz = g' (.) bits gives the same bf16 plane bit for bit
    gzp = split(gp * bits, False)
    out_m = ops.new_planes(B, L, Fn, False, dev)
    ops.conv3(gzp, bwd, (0, -d, -2 * d), 1, L, resid=gpp, resid_off=-d, out=out_m)
    torch.cuda.synchronize()
    assert torch.equal(out[0], out_m[0]), "masked-operand path differs from the materialised one"


@pytest.mark.parametrize("B,L,d", XF_CASES, ids=["-".join(map(str, c)) for c in XF_CASES])
def test_dilated_layer_weight_gradient_masked_operand(dev, B, L, d):
    """dW_j = sum a[t + j d]^T (g (.) bits)[t], db = sum g (.) bits with the bits applied to the
    staged tile in shared memory (bpb_wgrad3_args.gz_mask): bit-identical to the same launch on a
    materialised gz plane."""
    from bpreveal_b200 import ops
    Fn = 64
    g = torch.Generator().manual_seed(13)
    L_out = L - 2 * d
    a = split(torch.randn((B, L, Fn), generator=g).to(dev), False)
    gfull = torch.randn((B, L_out, Fn), generator=g).to(dev)
    bits = (torch.rand((B, L_out, Fn), generator=g) > 0.4).to(dev)
    res = []
    for masked in (True, False):
        gz = split(gfull if masked else gfull * bits, False)
        dw = torch.empty((3, Fn, Fn), dtype=torch.float32, device=dev)
        db = torch.empty((Fn,), dtype=torch.float32, device=dev)
        ws = torch.empty((ops.wgrad3_ws_bytes(B, L_out, Fn, False) // 4,), dtype=torch.float32,
                         device=dev)
        ops.wgrad3(a, gz, d, dw, db, ws, gz_mask=pack_mask(bits) if masked else None)
        torch.cuda.synchronize()
        res.append((dw, db))
    av, gv = val(a), val(split(gfull * bits, False))
    ref = torch.stack([torch.einsum("btc,btn->cn", av[:, j * d:j * d + L_out, :], gv)
                       for j in range(3)])
    assert Hh.rel_err(res[0][0].cpu().numpy(), ref.numpy()) < 2e-3
    assert Hh.rel_err(res[0][1].cpu().numpy(), gv.sum(dim=(0, 1)).numpy()) < 2e-3
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])


def test_masked_operand_backward_replays_bit_for_bit(dev):
    """C1's nine layer shapes at batch 64 through the masked-operand data gradient: 20 replays of
    a 10-launch CUDA graph must reproduce the first launch (the masker warps write the staged
    tile between the TMA load and the MMA: a missing fence or an early slot release shows here)."""
    from bpreveal_b200 import ops
    B, Fn, L = 64, 64, 3066
    g = torch.Generator().manual_seed(19)
    for i in range(9):
        d = 2 ** (i + 1)
        L_out = L - 2 * d
        gp = split(torch.randn((B, L_out, Fn), generator=g).to(dev), False)
        w = (torch.randn((3, Fn, Fn), generator=g) / 14).to(dev)
        _, bwd = prep_w(w, Fn, False, dev)
        gin = ops.new_planes(B, L, Fn, False, dev)
        mask_a = pack_mask((torch.rand((1, L_out, Fn), generator=g) > 0.5).to(dev)).expand(
            B, L_out, 1).contiguous()

        def f_bwd():
            ops.conv3(gp, bwd, (0, -d, -2 * d), 1, L, resid=gp, resid_off=-d, out=gin,
                      mask_a=mask_a)

        f_bwd()
        torch.cuda.synchronize()
        first = gin[0].clone()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(10):
                f_bwd()
        for _ in range(20):
            gr.replay()
            torch.cuda.synchronize()
            assert torch.equal(first, gin[0]), f"layer {i} (d={d}): replay differs"
        L = L_out


@pytest.mark.parametrize("name,cfg", [("C1", C1), ("C3R", C3R)])
def test_masked_operand_step_equals_materialised_step(dev, monkeypatch, name, cfg):
    """BPB_XF=1 (no gz planes: operands masked in shared memory by the data- and weight-gradient
    kernels) against the default backward on the full-size network: every gradient and both
    losses bit for bit, and the same parameters after two optimiser steps."""
    B = 4
    x = torch.tensor(O.synthetic_sequences(B, cfg["inputLength"], seed=43), device=dev)
    counts = [O.synthetic_profiles(B, 1000, t, seed=43 + i, meanReads=150.0)
              for i, t in enumerate(cfg["headTasks"])]
    counts = torch.tensor(np.concatenate(counts, axis=2), device=dev)
    res = {}
    for xf in ("0", "1"):
        monkeypatch.setenv("BPB_XF", xf)
        net, _ = build(cfg, False, dev, scale=1.0)
        net.enable_training()
        ws = net._workspace(B, "train")
        assert net._xf_backward(ws) == (xf == "1")
        ws["x"].copy_(x)
        ws["counts"].copy_(counts)
        net.forward_backward_into(ws)
        torch.cuda.synchronize()
        g = net.grads_flat.clone()
        losses = ws["losses"].clone()
        for _ in range(2):
            net.train_step(x, counts, lr=0.004)
        torch.cuda.synchronize()
        res[xf] = (g, losses, net.params.flat.clone())
    for a, b, what in zip(res["0"], res["1"], ("gradients", "losses", "parameters after 2 steps")):
        assert torch.equal(a, b), f"{name}: {what} differ between the two backward paths"


CG2_CASES = [  # (B, L_in, dilation): F = 512; odd and even counts of 128-row tiles, one tile
    (1, 132, 2), (1, 600, 2), (2, 644, 2), (3, 900, 128), (1, 1400, 512), (2, 2308, 64),
]


@pytest.mark.parametrize("B,L,d", CG2_CASES, ids=["-".join(map(str, c)) for c in CG2_CASES])
def test_cta_pair_launches_equal_single_cta_launches(dev, monkeypatch, B, L, d):
    """F = 512 training launches run as CTA pairs (cta_group::2: M = 256 MMAs, each CTA stages
    half of the weight tile); BPB_CG2=0 selects the single-CTA kernel.  Forward (output + ReLU
    bits) and data gradient (g and gz) must agree bit for bit -- every output element sums the
    same products in the same order -- including when the second half of the last pair tile lies
    beyond the sequence."""
    from bpreveal_b200 import ops
    Fn = 512
    g = torch.Generator().manual_seed(21)
    L_out = L - 2 * d
    x = split(torch.randn((B, L, Fn), generator=g).to(dev), False)
    gz = split(torch.randn((B, L_out, Fn), generator=g).to(dev), False)
    gp = split(torch.randn((B, L_out, Fn), generator=g).to(dev), False)
    w = (torch.randn((3, Fn, Fn), generator=g) / (3 * Fn) ** 0.5).to(dev)
    bias = (torch.randn((Fn,), generator=g) * 0.1).to(dev)
    fwd, bwd = prep_w(w, Fn, False, dev)
    mask_in = pack_mask((torch.rand((B, L, Fn), generator=g) > 0.5).to(dev))
    res = {}
    for cg2 in ("0", "1"):
        monkeypatch.setenv("BPB_CG2", cg2)
        out = ops.new_planes(B, L_out, Fn, False, dev)
        mask = torch.zeros((B, L_out, Fn // 64), dtype=torch.int64, device=dev)
        ops.conv3(x, fwd, (0, d, 2 * d), 0, L_out, bias=bias, resid=x, resid_off=d, out=out,
                  mask_out=mask)
        gin = ops.new_planes(B, L, Fn, False, dev)
        gz2 = ops.new_planes(B, L, Fn, False, dev)
        ops.conv3(gz, bwd, (0, -d, -2 * d), 1, L, resid=gp, resid_off=-d, out=gin, out2=gz2,
                  mask_in=mask_in)
        torch.cuda.synchronize()
        res[cg2] = (out[0], mask, gin[0], gz2[0])
    for a, b, what in zip(res["0"], res["1"], ("forward", "ReLU bits", "g", "gz")):
        assert torch.equal(a, b), f"{what} differs between CTA-pair and single-CTA launches"
    xe = val(x)
    ref = torch.relu(O._conv1d_cl(xe, w_eff(w, False), bias.double().cpu(), dilation=d)) + \
        xe[:, d:L - d, :]
    assert Hh.rel_err(val((res["1"][0], None)).numpy(), ref.numpy()) < 6e-3


def test_cta_pair_launches_replay_bit_for_bit(dev):
    """F = 512 forward and data gradient as CTA pairs, C4-like lengths at batch 4: 15 replays of
    an 8-launch CUDA graph must reproduce the first launch (remote mbarrier arrivals, multicast
    commits and the two producers of a stage: a missing cluster-scope ordering shows here)."""
    from bpreveal_b200 import ops
    B, Fn = 4, 512
    g = torch.Generator().manual_seed(23)
    for L, d in ((9262, 2), (6000, 256), (5200, 2048)):
        L_out = L - 2 * d
        x = split(torch.randn((B, L, Fn), generator=g).to(dev), False)
        w = (torch.randn((3, Fn, Fn), generator=g) / 40).to(dev)
        fwd, bwd = prep_w(w, Fn, False, dev)
        bias = torch.zeros((Fn,), device=dev)
        out = ops.new_planes(B, L_out, Fn, False, dev)
        mask = torch.zeros((B, L_out, Fn // 64), dtype=torch.int64, device=dev)
        gin = ops.new_planes(B, L, Fn, False, dev)
        gz2 = ops.new_planes(B, L, Fn, False, dev)
        mask_in = torch.randint(-2 ** 62, 2 ** 62, (B, L, Fn // 64), dtype=torch.int64, device=dev)

        def f_fwd():
            ops.conv3(x, fwd, (0, d, 2 * d), 0, L_out, bias=bias, resid=x, resid_off=d, out=out,
                      mask_out=mask)

        def f_bwd():
            ops.conv3(out, bwd, (0, -d, -2 * d), 1, L, resid=out, resid_off=-d, out=gin,
                      out2=gz2, mask_in=mask_in)

        for fn, res in ((f_fwd, lambda: (out[0], mask)), (f_bwd, lambda: (gin[0], gz2[0]))):
            fn()
            torch.cuda.synchronize()
            first = [t.clone() for t in res()]
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(8):
                    fn()
            for _ in range(15):
                gr.replay()
                torch.cuda.synchronize()
                for a_, b_ in zip(first, res()):
                    assert torch.equal(a_, b_), f"L={L} d={d}: replay differs"


def test_every_c1_tower_shape_many_launches_in_a_graph(dev):
    """Each of C1's nine layer shapes at batch 64, forward and backward, 30 replays of a
    10-launch CUDA graph; every replay must reproduce the first launch bit for bit.  This test
    found a real race in round 2: the forward at d = 64 (HALO staging, 4 box slots) released a
    box slot to the TMA producer right behind the ld.shared instructions that read the residual
    out of it, and ~1 launch in 100 added the residual of the tile four steps ahead
    (conv_core.cuh, "release")."""
    from bpreveal_b200 import ops
    B, Fn, L = 64, 64, 3066
    g = torch.Generator().manual_seed(9)
    for i in range(9):
        d = 2 ** (i + 1)
        L_out = L - 2 * d
        x = split(torch.randn((B, L, Fn), generator=g).to(dev), False)
        w = (torch.randn((3, Fn, Fn), generator=g) / 14).to(dev)
        fwd, bwd = prep_w(w, Fn, False, dev)
        out = ops.new_planes(B, L_out, Fn, False, dev)
        mask = torch.zeros((B, L_out, 1), dtype=torch.int64, device=dev)
        bias = torch.zeros((Fn,), device=dev)
        gin = ops.new_planes(B, L, Fn, False, dev)
        gz2 = ops.new_planes(B, L, Fn, False, dev)
        mask_in = pack_mask((torch.rand((1, L, Fn), generator=g) > 0.5).to(dev)).expand(
            B, L, 1).contiguous()

        def f_fwd():
            ops.conv3(x, fwd, (0, d, 2 * d), 0, L_out, bias=bias, resid=x, resid_off=d, out=out,
                      mask_out=mask)

        def f_bwd():
            ops.conv3(out, bwd, (0, -d, -2 * d), 1, L, resid=out, resid_off=-d, out=gin,
                      out2=gz2, mask_in=mask_in)

        for fn, res in ((f_fwd, lambda: (out[0], mask)), (f_bwd, lambda: (gin[0], gz2[0]))):
            fn()
            torch.cuda.synchronize()
            first = [t.clone() for t in res()]
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(10):
                    fn()
            for _ in range(30):  # 300 launches: the d = 64 race of round 1 hit ~1 launch in 100
                gr.replay()
                torch.cuda.synchronize()
                for a_, b_ in zip(first, res()):
                    assert torch.equal(a_, b_), f"layer {i} (d={d}): replay differs"
        L = L_out


# ------------------------------------------------------------------------------------------
# (iii) C4: F = 512, 11 layers (dilations to 2048), widths 25 / 75
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precise", [False, True])
def test_c4_shaped_forward_loss_gradients(dev, precise):
    cfg = C4S
    net, params = build(cfg, precise, dev, scale=1.0)
    net.enable_training()
    B = 1
    x = O.synthetic_sequences(B, cfg["inputLength"], seed=51)
    counts = O.synthetic_profiles(B, cfg["outputLength"], 2, seed=51, meanReads=80.0)
    pw, lam = [1.0], [3.0]
    net.profile_weight.copy_(torch.tensor(pw))
    net.lam.copy_(torch.tensor(lam))
    ref_l, ref_c = Hh.oracle_forward(params, x)
    tot, pl, cl, grads = Hh.oracle_loss_and_grads(params, x, counts, pw, lam)
    got_l, got_c = net.predict(torch.tensor(x, device=dev), batchSize=1)
    el, ec = Hh.rel_err(got_l.cpu().numpy(), ref_l), Hh.rel_err(got_c.cpu().numpy(), ref_c)
    ws = net._workspace(B, "train")
    ws["x"].copy_(torch.tensor(x, device=dev))
    ws["counts"].copy_(torch.tensor(counts, device=dev))
    net.forward_backward_into(ws)
    torch.cuda.synchronize()
    losses = ws["losses"].cpu().numpy()
    e_mnll = abs(losses[0] - pl[0]) / abs(pl[0])
    e_mse = abs(losses[1] - cl[0]) / abs(cl[0])
    ge = grad_errors(net.state(net.grads_flat), grads)
    got = dict(logits=el, logcounts=ec, mnll=e_mnll, wmse=e_mse)
    floor = {} if precise else bf16_floor(params, x, counts, pw, lam)
    record(f"C4s_fwd_bwd_{'precise' if precise else 'bf16'}", grad_max=max(ge.values()),
           grad_dil10=ge["dil10_w"], grad_conv1=ge["conv1_w"], **got,
           **{f"floor_{k}": v for k, v in floor.items()})
    if precise:
        # (Until the streamed-weight launches dropped their residual ring, the two-plane data
        # gradient at F = 512 did not fit shared memory and this path was forward-only.)
        assert el < 1e-4 and ec < 1e-4, (el, ec)
        assert e_mnll < 1e-3 and e_mse < 1e-3, (e_mnll, e_mse)
    else:
        gate_bf16("C4s", got, floor, dict(logits=6e-2, logcounts=5e-2, mnll=2e-2, wmse=2e-2))
    # K = 1536 per layer, 13 layers deep: sign flips of near-zero pre-activations are what is
    # left on the precise path; on the bf16 path it is the rounding of every stored plane
    # (measured: the last layer's weight gradient 3.5e-5, growing layer by layer down to 1.2e-2
    # for the input convolution, thirteen ReLU layers below the loss)
    gtol = 2e-2 if precise else 8e-2
    for k, e in ge.items():
        assert e < gtol, f"grad {k}: rel err {e}"


# ------------------------------------------------------------------------------------------
# (iv) C3: combined model, 4 heads x 2 tasks, 7/7, use-bias-counts on one head
# ------------------------------------------------------------------------------------------
def _c3_models(dev, precise):
    from bpreveal_b200 import models
    names = ["oct4", "sox2", "klf4", "nanog"]
    heads = [{"head-name": n, "num-tasks": 2, "profile-loss-weight": 1.0 + 0.25 * i,
              "counts-loss-weight": 10.0 + i, "use-bias-counts": i == 1}
             for i, n in enumerate(names)]
    spec_p = {"name": "simple", "types": ["linear", "sigmoid"]}
    spec_c = {"name": "simple", "types": ["linear"]}
    # frozen bias model: F = 16, nine layers, widths 5/5 -> 3052 bp input, cropped by 2 each side
    bcfg = cfg_of(16, 9, 5, 5, 1000, [2, 2, 2, 2])
    solo = models.soloModel(bcfg["inputLength"], 1000, 16, 9, 5, 5, heads, "solo",
                            precise=precise, device=dev, seed=1)
    bparams = Hh.oracle_params(bcfg, seed=77, scale=1.0)
    solo.net.load_state(bparams)
    tm = models.transformationModel(solo, spec_p, spec_c, heads)
    rng = np.random.default_rng(5)
    tp = O.init_transform_params(4, spec_p, spec_c)
    for k in tp:
        tp[k] = (tp[k] + rng.standard_normal(2) * 0.2).astype(np.float32)
    named = dict(tm.named_weights())
    for h, n in enumerate(names):
        for key, base in ((f"profile{h}_linear", f"regress_linear_profile_{n}_conv"),
                          (f"profile{h}_sigmoid_in", f"sigmoid_in_linear_profile_{n}_conv"),
                          (f"profile{h}_sigmoid_out", f"sigmoid_out_linear_profile_{n}_conv"),
                          (f"logcounts{h}_linear", f"regress_linear_logcounts_{n}_conv")):
            named[base + "/kernel"] = tp[key][0].reshape(1, 1, 1)
            named[base + "/bias"] = tp[key][1].reshape(1)
    tm.set_weights(named)
    comb, resid, _ = models.combinedModel(C3R["inputLength"], 1000, 64, 9, 7, 7, heads, tm,
                                          precise=precise, seed=2)
    rparams = Hh.oracle_params(C3R, seed=78, scale=1.0)
    resid.net.load_state(rparams)
    return comb, resid, heads, (bcfg, bparams, tp, spec_p, spec_c, rparams)


@pytest.mark.parametrize("precise", [False, True])
def test_c3_combined_forward_loss_and_residual_gradients(dev, precise):
    """models.py:331-387 + layers.py:52-75: the gradient reaches the residual tower through the
    added logits and, for the head with use-bias-counts, through d/dr log(e^b + e^r) =
    sigmoid(r - b); the frozen bias model gets none."""
    from bpreveal_b200 import models
    comb, resid, heads, (bcfg, bparams, tp, spec_p, spec_c, rparams) = _c3_models(dev, precise)
    B = 2
    x = O.synthetic_sequences(B, C3R["inputLength"], seed=61)
    counts = O.synthetic_profiles(B, 1000, 8, seed=61, meanReads=120.0)
    pw = [h["profile-loss-weight"] for h in heads]
    lam = [h["counts-loss-weight"] for h in heads]
    useBias = [bool(h["use-bias-counts"]) for h in heads]
    # oracle
    rp = O.to_torch(rparams, torch.float64, requires_grad=True)
    bp = O.to_torch(bparams, torch.float64)
    tpt = {k: torch.tensor(v, dtype=torch.float64) for k, v in tp.items()}
    profs, cnts, _ = O.combined_forward(torch.tensor(x, dtype=torch.float64), rp, bp, tpt, spec_p,
                                        spec_c, useBias, bcfg["inputLength"])
    ct = torch.tensor(counts, dtype=torch.float64)
    trueP = [ct[:, :, 2 * h:2 * h + 2] for h in range(4)]
    trueC = [torch.log(t.sum(dim=(1, 2))) for t in trueP]
    tot, pl, cl = O.total_loss(profs, cnts, trueP, trueC, pw, lam)
    tot.backward()
    ref_g = {k: v.grad.numpy() for k, v in rp.items()}
    # device
    outs = comb.predict(x, batch_size=2)
    el = Hh.rel_err(np.concatenate(outs[:4], axis=2),
                    torch.cat(profs, dim=2).detach().numpy())
    ec = Hh.rel_err(np.concatenate(outs[4:], axis=1), torch.cat(cnts, dim=1).detach().numpy())
    comb.compile(optimizer=models.Adam(learning_rate=0.0), loss_weights=pw + [1.0] * 4)
    comb.set_lambdas(lam)
    bias_before = [v.copy() for v in comb.bias.get_weights()]
    losses = comb._step(x, counts, train=True)
    losses = losses.cpu().numpy() if isinstance(losses, torch.Tensor) else np.asarray(losses)
    e_mn = max(abs(losses[h] - pl[h].item()) / abs(pl[h].item()) for h in range(4))
    e_ms = max(abs(losses[4 + h] - cl[h].item()) / abs(cl[h].item()) for h in range(4))
    e_loss = max(e_mn, e_ms)
    ge = grad_errors(resid.net.state(resid.net.grads_flat), ref_g)
    record(f"C3_combined_{'precise' if precise else 'bf16'}", logits=el, logcounts=ec,
           mnll=e_mn, wmse=e_ms, grad_max=max(ge.values()), grad_cnt1=ge["cnt1_w"],
           grad_conv1=ge["conv1_w"])
    if precise:
        assert el < 1e-4 and ec < 1e-4, (el, ec)
        assert e_loss < 1e-3, f"losses rel err {e_loss}"
    else:  # the residual tower's bf16 storage floor, as for C1 (two towers add up here)
        assert el < 2e-2 and ec < 2e-2, (el, ec)
        assert e_mn < 1e-2 and e_ms < 3e-2, (e_mn, e_ms)
    gtol = 2e-3 if precise else 5e-2
    for k, e in ge.items():
        assert e < gtol, f"residual grad {k}: rel err {e}"
    for u, v in zip(bias_before, comb.bias.get_weights()):
        assert np.array_equal(u, v)


def test_c3_combined_checkpoint_roundtrip(dev, tmp_path):
    """residual and bias head variables share Keras names
    (solo_profile_{head}, models.py:41-48); a saved combined model must bring both back."""
    from bpreveal_b200 import models
    comb, resid, heads, _ = _c3_models(dev, False)
    names = [k for k, _ in comb.named_weights()]
    assert len(set(names)) == len(names), "weight names collide"
    x = O.synthetic_sequences(2, C3R["inputLength"], seed=5)
    want = comb.predict(x)
    for ext in (".npz", ".keras", ".h5"):
        path = str(tmp_path / ("c" + ext))
        comb.save(path)
        back = models.loadModel(path, device=dev)
        assert back.name == "combined_model" and back.cropFlank == comb.cropFlank
        for u, v in zip(want, back.predict(x)):
            assert np.array_equal(u, v), ext


# ------------------------------------------------------------------------------------------
# (v) C5: PISA on the C1 model, 20 shuffled references, 3090-bp windows
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precise", [False, True])
def test_c5_pisa_20_shuffles_full_window(dev, precise):
    """interpretPisa.py:153-166 + shap.py:347-394, 612-639 at the BASELINE size: three query
    bases, 20 references each (seed 355687 permutations), attribution of the leftmost logit.
    Tolerance: max |phi - phi_oracle| <= 2e-3 (precise) / 5e-2 (bf16) of max |phi_oracle|;
    efficiency sum(phi) = v(x) - mean v(refs) on the device result to 2e-3 / 1e-1."""
    from bpreveal_b200 import interpret
    net, params = build(C1, precise, dev, scale=1.0)
    Q, n = 3, 20
    xs = O.synthetic_sequences(Q, C1["inputLength"], seed=71)
    p64 = O.to_torch(params, torch.float64)
    phis, vx, vr = interpret.pisa_batch(net, torch.tensor(xs, device=dev), headID=0, taskID=1,
                                        numShuffles=n)
    phis, vx, vr = phis.cpu().numpy(), vx.cpu().numpy(), vr.cpu().numpy()
    rf = C1["inputLength"] - 1000 + 1
    assert rf == 2091 and np.abs(phis[:, rf:]).max() == 0.0   # interpretPisa.py:45-48
    worst_phi = worst_eff = worst_v = 0.0
    for q in range(Q):
        refs = O.generate_shuffles(xs[q], n)
        phi, v, rv = O.deepshap(xs[q], refs, O.solo_shap_forward(p64),
                                lambda pr, ct: O.pisa_metric(pr, ct, 0, 1))
        scale = np.abs(phi).max()
        worst_phi = max(worst_phi, np.abs(phis[q] - phi).max() / scale)
        worst_v = max(worst_v, abs(vx[q] - v) / max(1.0, abs(v)),
                      np.abs(vr[q] - rv).max() / max(1.0, np.abs(rv).max()))
        rhs = vx[q] - vr[q].mean()
        worst_eff = max(worst_eff, abs(phis[q].sum() - rhs) / max(abs(rhs), scale))
    record(f"C5_pisa_{'precise' if precise else 'bf16'}", phi=worst_phi, efficiency=worst_eff,
           metric=worst_v)
    assert worst_phi < (2e-3 if precise else 5e-2), f"phi rel err {worst_phi}"
    # the sum runs over 2091 x 4 attributions, each carrying its own rounding
    assert worst_eff < (2e-3 if precise else 1e-1), f"efficiency gap {worst_eff}"
    assert worst_v < (1e-4 if precise else 3e-2), f"metric rel err {worst_v}"
