"""GPU parity of the whole network (forward, loss, gradients, Adam, DeepSHAP) against the CPU
oracle.  Tolerances: north_star states 1e-2 relative (bf16 path) and 1e-4 (fp32-accumulate
path) for logits / log-counts and 1e-3 for losses."""
import numpy as np
import pytest
import torch

from oracle import bpnet_oracle as O
from tests import helpers as Hh

pytestmark = pytest.mark.gpu

CASES = [
    ("tiny", Hh.tiny_cfg()),
    ("tiny2heads", Hh.tiny_cfg(headTasks=[1, 3], numFilters=16)),
    ("mid", Hh.mid_cfg()),
    ("mid128", Hh.mid_cfg(numFilters=128, numLayers=3, outputLength=150, headTasks=[2, 2])),
    # the wide regime of C4 in miniature: N tiles of 256 with streamed weights, width-75 heads
    ("wide256", Hh.mid_cfg(numFilters=256, numLayers=3, outputLength=60, outputFilterWidth=75)),
    # ragged everything: F = 24 (zero-padded to 64), even filter widths, an output length that is
    # not a multiple of the 128-row tile, a head of 9 tasks (beyond the fused loss kernel's 8, so
    # the step takes the separate loss / pack kernels) next to a single-task head
    ("ragged", Hh.mid_cfg(numFilters=24, numLayers=2, outputLength=131, inputFilterWidth=4,
                          outputFilterWidth=10, headTasks=[1, 9])),
    # no dilated layers at all: input convolution straight into the heads
    ("nolayers", Hh.mid_cfg(numFilters=16, numLayers=0, outputLength=40, inputFilterWidth=7,
                            outputFilterWidth=5, headTasks=[2])),
]


def build(cfg, precise, dev, seed=1234, scale=1.0):
    from bpreveal_b200.engine import NetConfig, SoloNet
    net = SoloNet(NetConfig(precise=precise, **cfg), dev)
    params = Hh.oracle_params(cfg, seed=seed, scale=scale)
    net.load_state(params)
    return net, params


@pytest.mark.parametrize("name,cfg", CASES)
@pytest.mark.parametrize("precise", [False, True])
def test_forward_parity(dev, name, cfg, precise):
    net, params = build(cfg, precise, dev)
    B = 5
    x = O.synthetic_sequences(B, cfg["inputLength"], seed=7)
    ref_l, ref_c = Hh.oracle_forward(params, x)
    got_l, got_c = net.predict(torch.tensor(x, device=dev), batchSize=3)
    tol = 1e-4 if precise else 1e-2
    el, ec = Hh.rel_err(got_l.cpu().numpy(), ref_l), Hh.rel_err(got_c.cpu().numpy(), ref_c)
    assert el < tol, f"logits rel err {el}"
    assert ec < tol, f"logcounts rel err {ec}"


@pytest.mark.parametrize("name,cfg", CASES)
@pytest.mark.parametrize("precise", [False, True])
def test_loss_and_gradients(dev, name, cfg, precise):
    net, params = build(cfg, precise, dev, scale=1.5)
    net.enable_training()
    B = 6
    H = len(cfg["headTasks"])
    T = sum(cfg["headTasks"])
    x = O.synthetic_sequences(B, cfg["inputLength"], seed=11)
    counts = O.synthetic_profiles(B, cfg["outputLength"], T, seed=11, meanReads=50.0)
    pw = [1.0 + 0.5 * h for h in range(H)]
    lam = [2.0 + h for h in range(H)]
    net.profile_weight.copy_(torch.tensor(pw))
    net.lam.copy_(torch.tensor(lam))
    tot, pl, cl, grads = Hh.oracle_loss_and_grads(params, x, counts, pw, lam)
    ws = net._workspace(B, "train")
    ws["x"].copy_(torch.tensor(x, device=dev))
    ws["counts"].copy_(torch.tensor(counts, device=dev))
    net.forward_backward_into(ws)
    torch.cuda.synchronize()
    losses = ws["losses"].cpu().numpy()
    ltol = 1e-3 if (precise or name.startswith("mid")) else 5e-3  # tiny nets: few terms per sum
    for h in range(H):
        assert abs(losses[h] - pl[h]) <= ltol * abs(pl[h]) + 1e-5, (h, losses[h], pl[h])
        assert abs(losses[H + h] - cl[h]) <= (1e-3 if precise else 3e-2) * abs(cl[h]) + 1e-5, \
            (h, losses[H + h], cl[h])
    got = net.state(net.grads_flat)
    # precise path: activations carry 16 mantissa bits (hi + lo bf16 planes), so pre-activations
    # within ~1e-5 of zero can land on the other side of the ReLU than in the fp64 oracle; each
    # such flip moves a gradient entry by one term.  K = 3 * 256 doubles the chance per element.
    gtol = (2e-3 if cfg["numFilters"] <= 128 else 5e-3) if precise else 8e-2
    gscale = max(np.abs(v).max() for v in grads.values())
    for k in grads:
        # relative to the tensor's own magnitude, floored by the global gradient scale (the profile
        # bias gradient of a single-task head is identically zero by softmax shift invariance)
        e = np.abs(got[k] - grads[k]).max() / max(np.abs(grads[k]).max(), 1e-3 * gscale)
        assert e < gtol, f"grad {k}: rel err {e}"


@pytest.mark.parametrize("precise", [False, True])
def test_adam_training_curve(dev, precise):
    """Five optimisation steps from the same init and batches follow the oracle's loss curve."""
    cfg = Hh.tiny_cfg(numFilters=16)
    net, params = build(cfg, precise, dev)
    B, T = 8, 2
    lr = 0.004
    p = O.to_torch(params, torch.float64, requires_grad=True)
    m = {k: torch.zeros_like(v) for k, v in p.items()}
    v_ = {k: torch.zeros_like(v) for k, v in p.items()}
    ref_curve, got_curve = [], []
    for step in range(1, 6):
        x = O.synthetic_sequences(B, cfg["inputLength"], seed=100 + step)
        counts = O.synthetic_profiles(B, cfg["outputLength"], T, seed=100 + step, meanReads=40.0)
        xt = torch.tensor(x, dtype=torch.float64)
        ct = torch.tensor(counts, dtype=torch.float64)
        profs, cnts = O.solo_forward(xt, p)
        tot, pl, cl = O.total_loss(profs, cnts, [ct], [torch.log(ct.sum(dim=(1, 2)))], [1.0],
                                   [1.0])
        for t in p.values():
            t.grad = None
        tot.backward()
        with torch.no_grad():
            for k in p:
                O.adam_step(p[k], p[k].grad, m[k], v_[k], step, lr)
        ref_curve.append(pl[0].item() + cl[0].item())
        losses = net.train_step(torch.tensor(x, device=dev), torch.tensor(counts, device=dev), lr)
        got_curve.append(float(losses.sum().item()))
    ref_curve, got_curve = np.array(ref_curve), np.array(got_curve)
    assert np.all(np.abs(got_curve - ref_curve) <= (1e-3 if precise else 5e-3) * np.abs(ref_curve)), \
        (got_curve, ref_curve)
    # parameters after 5 steps
    got = net.state()
    for k in params:
        e = Hh.rel_err(got[k], p[k].detach().numpy())
        assert e < (2e-3 if precise else 5e-2), f"param {k}: rel err {e}"


@pytest.mark.parametrize("name,cfg", [CASES[0], CASES[2]])
@pytest.mark.parametrize("precise", [False, True])
def test_pisa_shap_parity(dev, name, cfg, precise):
    """DeepSHAP PISA attributions (interpretPisa.py:153-166, shap.py:347-394) and the efficiency
    identity sum(phi) = v(x) - mean v(refs) (doc/text/pisa.rst:46-55)."""
    from bpreveal_b200 import interpret
    net, params = build(cfg, precise, dev, scale=2.0)
    Q, n = 3, 6
    xs = O.synthetic_sequences(Q, cfg["inputLength"], seed=21)
    p64 = O.to_torch(params, torch.float64)
    phis, vx, vr = interpret.pisa_batch(net, torch.tensor(xs, device=dev), headID=0, taskID=1,
                                        numShuffles=n)
    phis, vx, vr = phis.cpu().numpy(), vx.cpu().numpy(), vr.cpu().numpy()
    for q in range(Q):
        refs = O.generate_shuffles(xs[q], n)
        phi, val, rvals = O.deepshap(xs[q], refs, O.solo_shap_forward(p64),
                                     lambda pr, ct: O.pisa_metric(pr, ct, 0, 1))
        scale = np.abs(phi).max()
        tol = 2e-3 if precise else 5e-2
        assert np.abs(phis[q] - phi).max() <= tol * scale, \
            (q, np.abs(phis[q] - phi).max(), scale)
        # weights are scaled x2 here; the bf16 error of a tiny (8-filter) net grows with it
        assert abs(vx[q] - val) <= (1e-4 if precise else 3e-2) * max(1.0, abs(val))
        assert np.abs(vr[q] - rvals).max() <= (1e-4 if precise else 3e-2) * max(
            1.0, np.abs(rvals).max())
        # efficiency on the device result itself
        lhs = phis[q].sum()
        rhs = vx[q] - vr[q].mean()
        assert abs(lhs - rhs) <= (2e-3 if precise else 5e-2) * max(abs(rhs), scale), (lhs, rhs)


def test_pisa_receptive_field_crop_and_graph_are_exact(dev):
    """Evaluating PISA on the receptive field of output position 0 only, and replaying it as a
    CUDA graph, reproduces the eager whole-window computation (the convolutions are 'valid')."""
    from bpreveal_b200 import interpret
    cfg = Hh.mid_cfg()
    net, params = build(cfg, True, dev, scale=2.0)
    Q, n = 2, 5
    xs = torch.tensor(O.synthetic_sequences(2 * Q, cfg["inputLength"], seed=33), device=dev)
    full = interpret.pisa_batch(net, xs[:Q], 0, 1, n, use_graph=False, crop=False)
    fast = interpret.pisa_batch(net, xs[:Q], 0, 1, n)            # first call: eager warm-up
    again = interpret.pisa_batch(net, xs[Q:], 0, 1, n)           # graph replay on other inputs
    full2 = interpret.pisa_batch(net, xs[Q:], 0, 1, n, use_graph=False, crop=False)
    rf = cfg["inputLength"] - cfg["outputLength"] + 1
    for a, b in ((full, fast), (full2, again)):
        scale = a[0].abs().max().item()
        assert (a[0] - b[0]).abs().max().item() <= 1e-5 * scale
        assert b[0][:, rf:].abs().max().item() == 0.0
        assert torch.allclose(a[1], b[1], rtol=1e-5, atol=1e-6)
        assert torch.allclose(a[2], b[2], rtol=1e-5, atol=1e-6)
    # new weights invalidate the cached twin
    net.load_state(Hh.oracle_params(cfg, seed=99))
    other = interpret.pisa_batch(net, xs[:Q], 0, 1, n)
    ref = interpret.pisa_batch(net, xs[:Q], 0, 1, n, use_graph=False, crop=False)
    assert (other[0] - ref[0]).abs().max().item() <= 1e-5 * ref[0].abs().max().item()


@pytest.mark.parametrize("precise", [False, True])
def test_flat_hypothetical_scores_parity(dev, precise):
    """interpretFlat profile / counts metrics with the hypothetical-contribution combiner
    (interpretFlat.py:187-312, interpretUtils.py:1295-1312) against the oracle."""
    from bpreveal_b200 import interpret
    cfg = Hh.tiny_cfg(headTasks=[2, 1])
    net, params = build(cfg, precise, dev, scale=1.5)
    Q, n = 2, 6
    xs = O.synthetic_sequences(Q, cfg["inputLength"], seed=5)
    p64 = O.to_torch(params, torch.float64)
    xt = torch.tensor(xs, device=dev)
    hyp_p, vx_p, vr_p = interpret.flat_batch(net, xt, 0, n, kind="profile", taskIDs=[0, 1])
    hyp_c, vx_c, vr_c = interpret.flat_batch(net, xt, 1, n, kind="counts")
    tol = 3e-3 if precise else 6e-2
    for q in range(Q):
        refs = O.generate_shuffles(xs[q], n)
        ref_p, val_p, rv_p = O.deepshap(xs[q], refs, O.solo_shap_forward(p64),
                                        lambda pr, cn: O.profile_metric(pr, cn, 0, [0, 1], shap=True),
                                        hypothetical=True)
        ref_c, val_c, rv_c = O.deepshap(xs[q], refs, O.solo_shap_forward(p64),
                                        lambda pr, cn: O.counts_metric(pr, cn, 1), hypothetical=True)
        for got, ref in ((hyp_p[q], ref_p), (hyp_c[q], ref_c)):
            scale = np.abs(ref).max()
            assert np.abs(got.cpu().numpy() - ref).max() <= tol * scale
        assert abs(vx_p[q].item() - val_p) <= (1e-3 if precise else 5e-2) * max(1.0, abs(val_p))
        assert abs(vx_c[q].item() - val_c) <= (1e-4 if precise else 3e-2) * max(1.0, abs(val_c))
        assert np.abs(vr_c[q].cpu().numpy() - rv_c).max() <= (1e-4 if precise else 3e-2) * max(
            1.0, np.abs(rv_c).max())


@pytest.mark.parametrize("planes", [1, 2])
def test_fused_loss_pack_matches_separate_kernels(dev, planes):
    """bpb_loss_fwd_bwd_pack == bpb_loss_fwd_bwd + bpb_heads_pack_grad (bit-equal d8 and dlogits;
    batch sums in a different but fixed order), also on a second call (scratch state is reset)."""
    from bpreveal_b200 import ops
    g = torch.Generator().manual_seed(5)
    B, L, W, tasks = 5, 37, 7, [2, 3]
    T, H = sum(tasks), len(tasks)
    L_in, Cp = L + W - 1, ops.heads_cp(T, H)
    toff_h = [0, 2, 5]
    toff = torch.tensor(toff_h, dtype=torch.int32, device=dev)
    logits = torch.randn((B, L, T), generator=g).to(dev) * 2
    counts = torch.poisson(torch.rand((B, L, T), generator=g) * 4).to(dev)
    lc = torch.randn((B, H), generator=g).to(dev)
    pw = torch.tensor([1.0, 0.7], device=dev)
    lam = torch.tensor([2.0, 3.5], device=dev)
    losses_a = torch.zeros((2 * H,), device=dev)
    dl_a, dlc_a = torch.empty_like(logits), torch.empty_like(lc)
    ops.loss_fwd_bwd(toff, logits, lc, counts, pw, lam, losses_a, dl_a, dlc_a)
    d8_a = ops.heads_d8(B, L_in, W, Cp, planes, dev)
    dpb_a, dcb_a = torch.empty((T,), device=dev), torch.empty((H,), device=dev)
    ops.heads_pack_grad(dl_a, dlc_a, L_in, W, Cp, planes, d8_a, dpb=dpb_a, dcb=dcb_a)
    ws = ops.loss_pack_ws(B, T, H, dev)
    d8_b = ops.heads_d8(B, L_in, W, Cp, planes, dev)
    for _ in range(2):
        losses_b = torch.full((2 * H,), float("nan"), device=dev)
        dl_b, dlc_b = torch.empty_like(logits), torch.empty_like(lc)
        dpb_b, dcb_b = torch.empty((T,), device=dev), torch.empty((H,), device=dev)
        ops.loss_fwd_bwd_pack(toff, toff_h, logits, lc, counts, pw, lam, losses_b, L_in, W, Cp,
                              planes, d8_b, ws, dlogits=dl_b, dlogcounts=dlc_b, dpb=dpb_b, dcb=dcb_b)
        torch.cuda.synchronize()
        # (the two kernels add the softmax denominator in different orders)
        assert torch.allclose(dl_a, dl_b, rtol=1e-4, atol=1e-5 * float(dl_a.abs().max()))
        assert torch.equal(dlc_a, dlc_b)
        assert torch.allclose(d8_a.float(), d8_b.float(), rtol=1e-2, atol=1e-4 * float(dl_a.abs().max()))
        assert torch.allclose(losses_a, losses_b, rtol=1e-5)
        scale = dl_a.abs().sum(dim=(0, 1))
        assert ((dpb_a - dpb_b).abs() <= 1e-5 * scale).all()
        assert torch.allclose(dcb_a, dcb_b, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("precise", [False, True])
@pytest.mark.parametrize("case", [1, 4])
def test_fused_adam_refresh_matches_separate_kernels(dev, precise, case):
    """bpb_adam_prep_step == bpb_adam_step + the four bpb_prep_* kernels, bit for bit (parameters,
    moments and every bf16 operand layout), and it advances the device step counter.  Case 4
    (F = 256, 75-wide heads) keeps its heads' weights in per-channel-group images."""
    from bpreveal_b200 import ops
    cfg = dict(CASES[case][1])  # two heads | wide
    net, _ = build(cfg, precise, dev, scale=1.0)
    net.enable_training()
    g = torch.Generator().manual_seed(9)
    net.set_step(3, 0.004)
    net.grads_flat.copy_(torch.randn(net.grads_flat.shape, generator=g).to(dev) * 0.1)
    # padded entries of the bucket carry no gradient
    mask = torch.zeros_like(net.params.flat)
    for k, v in net.params.views(mask).items():
        if k == "dil_w":
            v[:, :, :net.cfg.numFilters, :net.cfg.numFilters] = 1
        elif k == "dil_b":
            v[:, :net.cfg.numFilters] = 1
        else:
            v.fill_(1)
    net.grads_flat.mul_(mask)
    net.adam_m.copy_(torch.randn(net.adam_m.shape, generator=g).to(dev) * 0.01 * mask)
    net.adam_v.copy_(torch.rand(net.adam_v.shape, generator=g).to(dev) * 0.01 * mask)
    names = ["w_fwd", "w_bwd", "w_in_op", "w_hd_op"] + (["w_hf_op"] if net.heads_tc else [])
    saved = {k: t.clone() for k, t in (("p", net.params.flat), ("m", net.adam_m), ("v", net.adam_v))}
    # separate kernels
    sl = net.step_and_lr.clone()
    sl[0] += 1
    ops.adam_step(net.params.flat, net.grads_flat, net.adam_m, net.adam_v, sl)
    net.refresh_weights()
    torch.cuda.synchronize()
    ref = {k: getattr(net, k).clone() for k in names}
    ref.update(p=net.params.flat.clone(), m=net.adam_m.clone(), v=net.adam_v.clone())
    # fused kernel from the same state, operand buffers scrambled where they will be rewritten
    net.params.flat.copy_(saved["p"])
    net.adam_m.copy_(saved["m"])
    net.adam_v.copy_(saved["v"])
    net.apply_adam()
    torch.cuda.synchronize()
    assert float(net.step_and_lr[0]) == 4.0
    assert torch.equal(net.params.flat, ref["p"])
    assert torch.equal(net.adam_m, ref["m"]) and torch.equal(net.adam_v, ref["v"])
    for k in names:
        assert torch.equal(getattr(net, k).view(torch.int16), ref[k].view(torch.int16)), k


def test_peer_exchange_kernel_with_one_rank_equals_plain_step(dev):
    """The fused all-reduce + Adam kernel (bpb_adam_prep_step_peer) on a world of one rank --
    bucket and flags in CUDA-IPC-capable memory, both flag barriers exercised over three steps --
    gives bit for bit the parameters of the plain step.  (Several ranks: tests/diag_peer.py under
    torchrun, which also checks that every rank ends with identical parameters.)"""
    from bpreveal_b200 import parallel
    cfg = dict(CASES[0][1])
    B = 4
    x = torch.tensor(O.synthetic_sequences(B, cfg["inputLength"], seed=3), device=dev)
    c = torch.tensor(O.synthetic_profiles(B, cfg["outputLength"], 2, seed=3, meanReads=30.0), device=dev)
    net_a, _ = build(cfg, False, dev)
    net_b, _ = build(cfg, False, dev)
    ex = parallel.PeerGradExchange(net_b.params.flat.numel(), dev, rank=0, world=1)
    try:
        net_b.enable_training(grads_flat=ex.grads)
        for _ in range(3):
            la = net_a.train_step(x, c, 0.004).clone()
            lb = net_b.train_step(x, c, 0.004, peers=ex).clone()
        torch.cuda.synchronize()
        assert torch.equal(la, lb)
        assert torch.equal(net_a.params.flat, net_b.params.flat)
        assert float(net_b.step_and_lr[0]) == 3.0
    finally:
        ex.close()


def test_training_from_host_batches_equals_device_batches(dev):
    """``train_step`` with pinned host tensors (two staging slots filled on a copy stream, one
    CUDA graph per slot reading the slot itself) against the same batches handed over as device
    tensors: identical parameters after six steps, with a validation pass and a device-input step
    interleaved (both must rebind the fixed input buffers while a slot is still bound)."""
    cfg = Hh.mid_cfg()
    B, T = 6, sum(cfg["headTasks"])
    xs = [O.synthetic_sequences(B, cfg["inputLength"], seed=70 + i) for i in range(3)]
    cs = [O.synthetic_profiles(B, cfg["outputLength"], T, seed=70 + i, meanReads=40.0)
          for i in range(3)]
    order = [0, 1, 2, 1, 0, 2]
    res = []
    for host in (False, True):
        net, _ = build(cfg, False, dev)
        net.enable_training()
        if host:
            xb = [torch.tensor(v).pin_memory() for v in xs]
            cb = [torch.tensor(v).pin_memory() for v in cs]
        else:
            xb = [torch.tensor(v, device=dev) for v in xs]
            cb = [torch.tensor(v, device=dev) for v in cs]
        losses = []
        for step, i in enumerate(order):
            if host and step == 3:
                # a device-input step and a validation pass between host-input steps
                ev = net.eval_losses(torch.tensor(xs[2], device=dev), torch.tensor(cs[2], device=dev))
                losses.append(ev.clone())
                out = net.train_step(torch.tensor(xs[i], device=dev), torch.tensor(cs[i], device=dev),
                                     0.004)
            else:
                if step == 3:
                    losses.append(net.eval_losses(xb[2], cb[2]).clone())
                out = net.train_step(xb[i], cb[i], 0.004)
            losses.append(out.clone())
        torch.cuda.synchronize()
        res.append((net.params.flat.clone(), torch.stack(losses)))
    dl = (res[0][1] - res[1][1]).abs().max(dim=1).values.tolist()
    dp = float((res[0][0] - res[1][0]).abs().max())
    train_rows = [r for r in range(len(dl)) if r != 3]  # row 3 is the validation pass
    assert all(dl[r] == 0.0 for r in train_rows), f"training losses differ (per row: {dl})"
    # (the validation kernel adds the batch with float atomics: last-bit jitter between runs)
    assert dl[3] <= 1e-5 * float(res[0][1][3].abs().max()), f"validation losses differ: {dl[3]}"
    assert torch.equal(res[0][0], res[1][0]), f"parameters differ by {dp}"
