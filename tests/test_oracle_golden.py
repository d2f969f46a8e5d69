"""CPU: the oracle against the committed fixtures under tests/golden/.

``reference_answers.npz`` holds outputs of the REFERENCE'S OWN code (imported from
/root/reference/pkg in the build container by tests/golden/make_golden.py, and libslide.c compiled
from the reference sources); these pin the oracle.  ``oracle_vectors.npz`` are regression pins of
the oracle's fp64 tensor arithmetic (the reference delegates that to TensorFlow, which cannot run
offline: "parity unpinned") plus mathematical identities that hold for any correct
restatement (fp64 gradient check, DeepSHAP efficiency).
"""
import os

import numpy as np
import pytest
import torch

from oracle import bpnet_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(GOLD, "reference_answers.npz"))


@pytest.fixture(scope="module")
def vec():
    return np.load(os.path.join(GOLD, "oracle_vectors.npz"))


# ------------------------------------------------------------------ pinned by the reference
def test_length_difference_matches_lengthcalc(ref):
    """lengthCalc.py:71-123 on a grid of (layers, input width, output width)."""
    for (L, wi, wo), want in zip(ref["lengthcalc_grid"], ref["lengthcalc_diff"]):
        assert O.length_difference(int(L), int(wi), int(wo)) == int(want)
    # the reference's own worked example (doc/demos/osknExample.ipynb cell 88): 3056 -> 1000 at 9/7/7
    assert 1000 + O.length_difference(9, 7, 7) == 3056
    # BASELINE configs[0]: 3090 -> 1000 at 9 layers, 25 / 23
    assert 1000 + O.length_difference(9, 25, 23) == 3090


def test_one_hot_encode_matches_utils(ref):
    """utils.py:420-486 incl. lower case and N -> all-zero rows."""
    for i in range(3):
        seq = str(ref[f"ohe_seq{i}"])
        got = O.one_hot_encode(seq, allowN=True)
        assert got.dtype == np.uint8
        assert np.array_equal(got, ref[f"ohe_out{i}"])
    with pytest.raises(AssertionError):
        O.one_hot_encode("ACGN", allowN=False)


def test_logits_to_profile_matches_utils(ref):
    got = O.logits_to_profile(ref["l2p_logits"], float(ref["l2p_logcounts"]))
    np.testing.assert_allclose(got, ref["l2p_profile"], rtol=2e-6, atol=0)


def test_shuffles_match_shap_batcher(ref):
    """interpretUtils.py:1216-1222, kmer 1: bit-exact permutations (same generator, same calls)."""
    x = ref["shuffle_x"]
    want = ref["shuffle_refs"]
    got = O.generate_shuffles(x, want.shape[0])
    assert np.array_equal(got, want)
    idx = O.shuffle_indices(x.shape[0], want.shape[0])
    assert np.array_equal(x[idx], want)
    # and the product's host-side permutation table (bpreveal_b200/interpret.py)
    from bpreveal_b200.interpret import shuffle_indices
    assert np.array_equal(shuffle_indices(x.shape[0], want.shape[0]), idx)


def test_combiners_match_reference(ref):
    """shap.py:26-30 (standard) and interpretUtils.py:1295-1312 (hypothetical)."""
    x = ref["shuffle_x"].astype(np.float64)
    refs = ref["shuffle_refs"].astype(np.float64)
    mult = ref["comb_mult"]
    std = (mult * (x[None] - refs)).mean(axis=0)
    np.testing.assert_allclose(std, ref["comb_standard"], rtol=1e-12, atol=1e-14)
    hyp = np.zeros_like(x)
    for i in range(4):
        e = np.zeros((1, 1, 4))
        e[0, 0, i] = 1.0
        hyp[:, i] = ((e - refs) * mult).sum(axis=2).mean(axis=0)
    np.testing.assert_allclose(hyp, ref["comb_hypothetical"], rtol=1e-12, atol=1e-14)


def test_slide_matches_libslide(ref):
    """libslide.c:44-91 (outputs of the library compiled from the reference sources)."""
    dst = np.zeros_like(ref["slide_dst"])
    O.slide(ref["slide_src"], dst, ref["slide_rows"], ref["slide_cols"])
    assert np.array_equal(dst, ref["slide_dst"])
    dstc = np.zeros_like(ref["slide_dstc"])
    O.slide(ref["slide_srcc"].astype(np.float32), dstc, ref["slide_rows"], ref["slide_cols"])
    assert np.array_equal(dstc, ref["slide_dstc"])


def test_slide_against_compiled_reference_if_present(ref):
    """When oracle/_ref/libslide.so travelled with the repo, run the reference C code live."""
    import ctypes
    so = os.path.join(os.path.dirname(GOLD), "..", "oracle", "_ref", "libslide.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libslide.so not built")
    lib = ctypes.CDLL(so)
    rng = np.random.default_rng(9)
    rows, scols, dcols, depth = 33, 120, 101, 2
    src = rng.standard_normal((rows, scols, depth)).astype(np.float32)
    ri = rng.permutation(rows).astype(np.int32)
    ci = rng.integers(0, scols - dcols + 1, size=rows).astype(np.int32)
    dst = np.zeros((rows, dcols, depth), dtype=np.float32)
    P = ctypes.c_void_p
    lib.slide(P(src.ctypes.data), P(dst.ctypes.data), ctypes.c_int(rows), ctypes.c_int(scols),
              ctypes.c_int(dcols), ctypes.c_int(depth), P(ri.ctypes.data), P(ci.ctypes.data))
    want = O.slide(src, np.zeros_like(dst), ri, ci)
    assert np.array_equal(dst, want)


# ------------------------------------------------------------------ regression pins + identities
def _params(vec, prefix):
    return {k[len(prefix):]: vec[k] for k in vec.files if k.startswith(prefix)}


def test_forward_loss_gradients_regression(vec):
    params = _params(vec, "p_")
    p = O.to_torch(params, torch.float64, requires_grad=True)
    xt = torch.tensor(vec["x"], dtype=torch.float64)
    ct = torch.tensor(vec["counts"], dtype=torch.float64)
    profs, cnts = O.solo_forward(xt, p)
    np.testing.assert_allclose(torch.cat(profs, dim=2).detach().numpy(), vec["logits"],
                               rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(torch.cat(cnts, dim=1).detach().numpy(), vec["logcounts"],
                               rtol=1e-10, atol=1e-12)
    trueP = [ct[:, :, 0:2], ct[:, :, 2:3]]
    trueC = [torch.log(t.sum(dim=(1, 2))) for t in trueP]
    tot, pl, cl = O.total_loss(profs, cnts, trueP, trueC, [1.0, 1.5], [2.0, 3.0])
    got = np.array([v.item() for v in pl] + [v.item() for v in cl] + [tot.item()])
    np.testing.assert_allclose(got, vec["losses"], rtol=1e-10)
    tot.backward()
    for k, v in p.items():
        np.testing.assert_allclose(v.grad.numpy(), vec[f"g_{k}"], rtol=1e-8, atol=1e-12)


def test_multinomial_nll_definition():
    """losses.py:15-39 == -mean_b log Multinomial(k_b; N_b, softmax(l_b)), checked against
    scipy's multinomial log-pmf on integer counts."""
    from scipy.stats import multinomial
    rng = np.random.default_rng(0)
    B, L, T = 3, 7, 2
    logits = rng.standard_normal((B, L, T))
    counts = rng.poisson(3.0, size=(B, L, T)).astype(np.float64)
    want = 0.0
    for b in range(B):
        lg = logits[b].reshape(-1)
        pr = np.exp(lg - lg.max())
        pr /= pr.sum()
        want -= multinomial.logpmf(counts[b].reshape(-1), n=counts[b].sum(), p=pr)
    want /= B
    got = O.multinomial_nll(torch.tensor(counts), torch.tensor(logits)).item()
    assert abs(got - want) < 1e-9 * abs(want)


def test_fp64_gradient_check(vec):
    """Central differences through the whole objective confirm the analytic gradients."""
    params = _params(vec, "p_")
    xt = torch.tensor(vec["x"], dtype=torch.float64)
    ct = torch.tensor(vec["counts"], dtype=torch.float64)
    trueP = [ct[:, :, 0:2], ct[:, :, 2:3]]
    trueC = [torch.log(t.sum(dim=(1, 2))) for t in trueP]

    def objective(pp):
        profs, cnts = O.solo_forward(xt, pp)
        return O.total_loss(profs, cnts, trueP, trueC, [1.0, 1.5], [2.0, 3.0])[0]

    rng = np.random.default_rng(4)
    for name in ("conv1_w", "dil1_w", "dil2_b", "prof0_w", "cnt1_w"):
        g = vec[f"g_{name}"]
        for _ in range(3):
            idx = tuple(int(rng.integers(0, s)) for s in g.shape)
            eps = 1e-6
            vals = []
            for sgn in (+1, -1):
                q = {k: v.copy() for k, v in params.items()}
                q[name][idx] = q[name][idx] + sgn * eps  # fp32 storage; delta is recomputed below
                vals.append((objective(O.to_torch(q, torch.float64)).item(), float(q[name][idx])))
            num = (vals[0][0] - vals[1][0]) / (vals[0][1] - vals[1][1])
            assert abs(num - g[idx]) <= 2e-4 * max(1.0, abs(g[idx])), (name, idx, num, g[idx])


def test_adam_regression(vec):
    """Five Keras-semantics Adam steps (eps outside the bias-corrected sqrt)."""
    params = _params(vec, "p_")
    p = O.to_torch(params, torch.float64, requires_grad=True)
    m = {k: torch.zeros_like(v) for k, v in p.items()}
    v_ = {k: torch.zeros_like(v) for k, v in p.items()}
    xt = torch.tensor(vec["x"], dtype=torch.float64)
    ct = torch.tensor(vec["counts"], dtype=torch.float64)
    trueP = [ct[:, :, 0:2], ct[:, :, 2:3]]
    trueC = [torch.log(t.sum(dim=(1, 2))) for t in trueP]
    curve = []
    for step in range(1, 6):
        profs, cnts = O.solo_forward(xt, p)
        tot, _, _ = O.total_loss(profs, cnts, trueP, trueC, [1.0, 1.5], [2.0, 3.0])
        for t in p.values():
            t.grad = None
        tot.backward()
        with torch.no_grad():
            for k in p:
                O.adam_step(p[k], p[k].grad, m[k], v_[k], step, 0.004)
        curve.append(tot.item())
    np.testing.assert_allclose(curve, vec["adam_curve"], rtol=1e-9)
    assert curve[-1] < curve[0]
    np.testing.assert_allclose(p["conv1_w"].detach().numpy(), vec["adam_conv1_w"], rtol=1e-8,
                               atol=1e-10)
    # first step of Adam moves every weight with a non-zero gradient by ~lr (sign update)
    p1 = torch.zeros(4, dtype=torch.float64)
    g1 = torch.tensor([1.0, -2.0, 1e-3, 0.0], dtype=torch.float64)
    O.adam_step(p1, g1, torch.zeros(4, dtype=torch.float64), torch.zeros(4, dtype=torch.float64),
                1, 0.01)
    # closed form of step 1: -lr * sqrt(1-b2) * g / (sqrt(1-b2)*|g| + eps): eps is NOT bias corrected
    s = np.sqrt(1 - 0.999)
    want = [-0.01 * s * g / (s * abs(g) + 1e-7) if g else 0.0 for g in (1.0, -2.0, 1e-3, 0.0)]
    np.testing.assert_allclose(p1.numpy(), want, rtol=1e-9, atol=1e-15)
    assert abs(p1[2].item() + 0.01) > 2e-5  # visibly different from torch.optim.Adam's placement


def test_deepshap_regression_and_efficiency(vec):
    """shap.py:347-394 + 612-639; efficiency sum(phi) = v(x) - mean v(refs)
    (doc/text/pisa.rst:46-55) and zero attribution outside the receptive field."""
    params = _params(vec, "p_")
    p64 = O.to_torch(params, torch.float64)
    x = vec["x"]
    refs = O.generate_shuffles(x[0], 6)
    phi, val, rvals = O.deepshap(x[0], refs, O.solo_shap_forward(p64),
                                 lambda pr, cn: O.pisa_metric(pr, cn, 0, 1))
    np.testing.assert_allclose(phi, vec["pisa_phi"], rtol=1e-9, atol=1e-13)
    assert abs(phi.sum() - (val - rvals.mean())) < 1e-10
    rf = x.shape[1] - 20 + 1  # receptive field of output position 0
    assert np.all(phi[rf:] == 0)
    phi_h, _, _ = O.deepshap(x[0], refs, O.solo_shap_forward(p64),
                             lambda pr, cn: O.profile_metric(pr, cn, 0, [0, 1], shap=True),
                             hypothetical=True)
    np.testing.assert_allclose(phi_h, vec["flat_profile_hyp"], rtol=1e-9, atol=1e-13)
    phi_c, _, _ = O.deepshap(x[0], refs, O.solo_shap_forward(p64),
                             lambda pr, cn: O.counts_metric(pr, cn, 1), hypothetical=True)
    np.testing.assert_allclose(phi_c, vec["flat_counts_hyp"], rtol=1e-9, atol=1e-13)
    # the hypothetical scores projected on the observed base equal the actual contributions
    phi_c_act, val_c, rv_c = O.deepshap(x[0], refs, O.solo_shap_forward(p64),
                                        lambda pr, cn: O.counts_metric(pr, cn, 1))
    np.testing.assert_allclose((phi_c * x[0]).sum(), phi_c_act.sum(), rtol=1e-9)
    assert abs(phi_c_act.sum() - (val_c - rv_c.mean())) < 1e-10


def test_combined_model_regression(vec):
    """models.py:268-387 (+ layers.py:52-75 fp64 CountsLogSumExp, use-bias-counts true/false)."""
    p64 = O.to_torch(_params(vec, "p_"), torch.float64)
    rp = O.to_torch(_params(vec, "r_"), torch.float64)
    tp = {k[2:]: torch.tensor(vec[k], dtype=torch.float64) for k in vec.files
          if k.startswith("t_")}
    spec_p = {"name": "simple", "types": ["linear", "sigmoid"]}
    spec_c = {"name": "simple", "types": ["linear"]}
    xr = torch.tensor(vec["xr"], dtype=torch.float64)
    profs, cnts, _ = O.combined_forward(xr, rp, p64, tp, spec_p, spec_c, [True, False], 52)
    np.testing.assert_allclose(torch.cat(profs, dim=2).numpy(), vec["comb_logits"], rtol=1e-10,
                               atol=1e-12)
    np.testing.assert_allclose(torch.cat(cnts, dim=1).numpy(), vec["comb_logcounts"], rtol=1e-10,
                               atol=1e-12)
    # head 1 has use-bias-counts false: its counts are the residual model's alone
    _, rc = O.solo_forward(xr, rp)
    np.testing.assert_allclose(cnts[1].numpy(), rc[1].numpy(), rtol=1e-12)
