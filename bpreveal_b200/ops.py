"""Python-side launchers for the C-ABI kernels (device pointers come from torch tensors).

A "plane set" is a tuple ``(hi, lo)`` of bf16 tensors ``[B, L, F]``; ``lo`` is ``None`` on the bf16
path.  Nothing here computes on the host: every function validates shapes and calls into
``libbpreveal_b200.so`` on the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import Conv3Args, HeadsDgradArgs, InputConvArgs, Wgrad3Args, check


def _p(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "native ops need contiguous CUDA tensors"
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _planes(ps):
    if ps is None:
        return None, None
    if isinstance(ps, torch.Tensor):
        return ps, None
    return ps[0], ps[1]


def new_planes(B, L, F, precise, device):
    hi = torch.empty((B, L, F), dtype=torch.bfloat16, device=device)
    lo = torch.empty((B, L, F), dtype=torch.bfloat16, device=device) if precise else None
    return (hi, lo)


def planes_to_float(ps):
    hi, lo = _planes(ps)
    v = hi.float()
    if lo is not None:
        v = v + lo.float()
    return v


def device_info():
    lib = _lib.load()
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    check(lib.bpb_device_info(C.byref(a), C.byref(b), C.byref(c)), "bpb_device_info")
    return a.value, b.value, c.value


def conv3(inp, w, tap_off, mode, L_out, bias=None, resid=None, resid_off=0, out=None, out2=None,
          mask_out=None, mask_in=None, mult=None, zx_in=None, zx_div=1, z_out=None, mask_a=None):
    """bpb_conv3_tc.  ``w`` is the prepared operand tensor [planes, 3, F, F] (bf16)."""
    lib = _lib.load()
    in_hi, in_lo = _planes(inp)
    B, L_in, F = in_hi.shape
    a = Conv3Args()
    a.B, a.L_in, a.L_out, a.F = B, L_in, L_out, F
    a.tap_off[0], a.tap_off[1], a.tap_off[2] = tap_off
    a.mode = mode
    a.in_hi, a.in_lo = _p(in_hi), _p(in_lo)
    assert w.dtype == torch.bfloat16 and w.shape[-3:] == (3, F, F)
    a.w_hi = _p(w)
    a.w_lo = (w.data_ptr() + 3 * F * F * 2) if in_lo is not None else None
    if in_lo is not None:
        assert w.shape[0] == 2, "precise path needs hi and lo weight planes"
    a.bias = _p(bias)
    r_hi, r_lo = _planes(resid)
    a.resid_hi, a.resid_lo = _p(r_hi), _p(r_lo)
    a.L_res = r_hi.shape[1] if r_hi is not None else 0
    a.resid_off = resid_off
    o_hi, o_lo = _planes(out)
    a.out_hi, a.out_lo = _p(o_hi), _p(o_lo)
    p_hi, p_lo = _planes(out2)
    a.out2_hi, a.out2_lo = _p(p_hi), _p(p_lo)
    for t in (o_hi, o_lo, p_hi, p_lo):
        assert t is None or tuple(t.shape) == (B, L_out, F)
    a.mask_out, a.mask_in = _p(mask_out), _p(mask_in)
    m_hi, m_lo = _planes(mult)
    a.mult_hi, a.mult_lo = _p(m_hi), _p(m_lo)
    a.zx_in, a.zx_div, a.z_out = _p(zx_in), zx_div, _p(z_out)
    a.mask_a = _p(mask_a)
    if mask_a is not None:
        assert tuple(mask_a.shape) == (B, L_in, F // 64) and mask_a.dtype == torch.int64
    check(lib.bpb_conv3_tc(C.byref(a), _stream()), "bpb_conv3_tc")


def wgrad3_ws_bytes(B, L_out, F, precise):
    n = _lib.load().bpb_wgrad3_ws_bytes(B, L_out, F, 1 if precise else 0)
    if n < 0:
        check(1, "bpb_wgrad3_ws_bytes")
    return n


def wgrad3(act, gz, dil, dw, db, ws, gz_mask=None):
    lib = _lib.load()
    a_hi, a_lo = _planes(act)
    g_hi, g_lo = _planes(gz)
    B, L_in, F = a_hi.shape
    a = Wgrad3Args()
    a.B, a.L_in, a.L_out, a.F, a.dil = B, L_in, g_hi.shape[1], F, dil
    a.act_hi, a.act_lo, a.gz_hi, a.gz_lo = _p(a_hi), _p(a_lo), _p(g_hi), _p(g_lo)
    assert dw.dtype == torch.float32 and dw.numel() == 3 * F * F
    a.dw, a.db = _p(dw), _p(db)
    a.ws, a.ws_bytes = _p(ws), ws.numel() * ws.element_size()
    if gz_mask is not None:
        assert tuple(gz_mask.shape) == (B, g_hi.shape[1], F // 64) and gz_mask.dtype == torch.int64
    a.gz_mask = _p(gz_mask)
    check(lib.bpb_wgrad3_tc(C.byref(a), _stream()), "bpb_wgrad3_tc")


def new_x8(B, L, device):
    """bf16 [B, L, 8] view of a buffer with the spare tail the STRIP-staged kernels read into."""
    n = _lib.load().bpb_x8_elems(B, L)
    flat = torch.zeros((n,), dtype=torch.bfloat16, device=device)
    return flat[:B * L * 8].view(B, L, 8)


def onehot_expand(x, x8):
    """u8 one-hot [B, L, 4] -> bf16 [B, L, 8] (the tcgen05 operand of the input convolution)."""
    B, L, nb = x.shape
    assert nb == 4 and x.dtype == torch.uint8 and tuple(x8.shape) == (B, L, 8)
    check(_lib.load().bpb_onehot_expand(B, L, _p(x), _p(x8), _stream()), "bpb_onehot_expand")


def input_conv_wop(W, F, planes, device):
    n = _lib.load().bpb_input_conv_wop_elems(W, F, planes)
    return torch.zeros((n,), dtype=torch.bfloat16, device=device)


def prep_input_conv_weights(w, F, w_op, planes):
    W, _, Fw = w.shape
    check(_lib.load().bpb_prep_input_conv_weights(W, Fw, F, _p(w), _p(w_op), planes, _stream()),
          "bpb_prep_input_conv_weights")


def input_conv_fwd(x8, w_op, w_planes, W, Fw, bias, F, out, mask_out=None, z_out=None, zx_in=None,
                   zx_div=1, mult=None):
    lib = _lib.load()
    B, L_in, c8 = x8.shape
    assert c8 == 8 and x8.dtype == torch.bfloat16
    a = InputConvArgs()
    a.B, a.L_in, a.L_out, a.W, a.Fw, a.F = B, L_in, L_in - W + 1, W, Fw, F
    a.x8, a.w_op, a.w_planes, a.bias = _p(x8), _p(w_op), w_planes, _p(bias)
    o_hi, o_lo = _planes(out)
    a.out_hi, a.out_lo = _p(o_hi), _p(o_lo)
    a.mask_out, a.z_out, a.zx_in, a.zx_div = _p(mask_out), _p(z_out), _p(zx_in), zx_div
    m_hi, m_lo = _planes(mult)
    a.mult_hi, a.mult_lo = _p(m_hi), _p(m_lo)
    check(lib.bpb_input_conv_fwd(C.byref(a), _stream()), "bpb_input_conv_fwd")


def _fill_wgrad3(a, layer):
    """(act planes, gz planes, dilation, dw, db[, gz_mask]) -> bpb_wgrad3_args.  With ``gz_mask``
    ([B, L_out, 1] int64 ReLU bits) the gz planes hold the UNMASKED gradient and the kernel masks
    the staged operand in shared memory (bf16 path, F = 64)."""
    act, gz, dil, dw, db = layer[:5]
    gz_mask = layer[5] if len(layer) > 5 else None
    a_hi, a_lo = _planes(act)
    g_hi, g_lo = _planes(gz)
    B, L_in, F = a_hi.shape
    a.B, a.L_in, a.L_out, a.F, a.dil = B, L_in, g_hi.shape[1], F, dil
    a.act_hi, a.act_lo, a.gz_hi, a.gz_lo = _p(a_hi), _p(a_lo), _p(g_hi), _p(g_lo)
    a.dw, a.db = _p(dw), _p(db)
    if gz_mask is not None:
        assert tuple(gz_mask.shape) == (B, g_hi.shape[1], F // 64) and gz_mask.dtype == torch.int64
    a.gz_mask = _p(gz_mask)


def wgrad3_multi_ws_bytes(B, L_outs, F):
    arr = (C.c_int * len(L_outs))(*L_outs)
    n = _lib.load().bpb_wgrad3_multi_ws_bytes(len(L_outs), B, arr, F)
    if n < 0:
        check(1, "bpb_wgrad3_multi_ws_bytes")
    return n


def wgrad3_multi(layers, ws):
    """layers: list of (act planes, gz planes, dilation, dw, db) -- every tower layer of a step in
    one launch (bpb_wgrad3_multi)."""
    lib = _lib.load()
    arr = (Wgrad3Args * len(layers))()
    for a, layer in zip(arr, layers):
        _fill_wgrad3(a, layer)
    check(lib.bpb_wgrad3_multi(arr, len(layers), _p(ws), ws.numel() * ws.element_size(),
                               _stream()), "bpb_wgrad3_multi")


def wgrad_tower_input_ws_bytes(B, L_outs, F, L_out0, W, L_last=0, W_out=0):
    """Workspace of ``wgrad_tower_input`` (W_out > 0: with the heads job); -1 when the jobs cannot
    share a launch."""
    arr = (C.c_int * len(L_outs))(*L_outs)
    return _lib.load().bpb_wgrad_tower_input_ws_bytes(len(L_outs), B, arr, F, L_out0, W, L_last,
                                                      W_out)


def wgrad_tower_input(layers, x8, gz0, W, Fw, dw1, db1, ws, heads=None, gz0_mask=None):
    """Every tower weight gradient of a step, the input convolution's and (``heads`` =
    (act_last planes, d8, W_out, T, H, dpw, dcw)) the heads' in one launch (bf16 path).  layers as
    in ``wgrad3_multi``; gz0 = gradient planes at the input conv's output (with ``gz0_mask``: the
    unmasked gradient and the input convolution's ReLU bits)."""
    lib = _lib.load()
    arr = (Wgrad3Args * len(layers))()
    for a, layer in zip(arr, layers):
        _fill_wgrad3(a, layer)
    g0_hi, _ = _planes(gz0)
    if heads is not None:
        act_last, d8, W_out, T, H, dpw, dcw = heads
        al_hi, _ = _planes(act_last)
        hargs = (al_hi.shape[1], W_out, T, H, _p(al_hi), _p(d8), _p(dpw), _p(dcw))
    else:
        hargs = (0, 0, 0, 0, None, None, None, None)
    check(lib.bpb_wgrad_tower_input(arr, len(layers), x8.shape[1], g0_hi.shape[1], W, Fw, _p(x8),
                                    _p(g0_hi), _p(gz0_mask), _p(dw1), _p(db1), *hargs, _p(ws),
                                    ws.numel() * ws.element_size(), _stream()),
          "bpb_wgrad_tower_input")


def input_conv_wgrad_ws_bytes(B, L_out, W, F):
    n = _lib.load().bpb_input_conv_wgrad_ws_bytes(B, L_out, W, F)
    if n < 0:
        check(1, "bpb_input_conv_wgrad_ws_bytes")
    return n


def input_conv_wgrad(x8, gz, W, Fw, dw, db, ws):
    lib = _lib.load()
    g_hi, g_lo = _planes(gz)
    B, L_out, F = g_hi.shape
    check(lib.bpb_input_conv_wgrad(B, x8.shape[1], L_out, W, Fw, F, _p(x8), _p(g_hi), _p(g_lo),
                                   _p(dw), _p(db), _p(ws), ws.numel() * ws.element_size(),
                                   _stream()), "bpb_input_conv_wgrad")


def input_conv_dgrad(gz, w, L_in, dx):
    lib = _lib.load()
    g_hi, g_lo = _planes(gz)
    B, L_out, F = g_hi.shape
    W, _, Fw = w.shape
    check(lib.bpb_input_conv_dgrad(B, L_in, L_out, W, Fw, F, _p(g_hi), _p(g_lo), _p(w), _p(dx),
                                   _stream()), "bpb_input_conv_dgrad")


def input_conv_dgrad_tc_supported(W, F, planes):
    return _lib.load().bpb_input_conv_dgrad_tc_supported(W, F, planes) == 0


def input_dgrad_wop(W, F, planes, device):
    n = _lib.load().bpb_input_dgrad_wop_elems(W, F, planes)
    return torch.zeros((n,), dtype=torch.bfloat16, device=device)


def prep_input_dgrad_weights(w, F, w_op, planes):
    W, _, Fw = w.shape
    check(_lib.load().bpb_prep_input_dgrad_weights(W, Fw, F, _p(w), _p(w_op), planes, _stream()),
          "bpb_prep_input_dgrad_weights")


def input_conv_dgrad_tc(gz, w_op, W, dx):
    """dx (B, L_in, 4) fp32 <- gz planes (B, L_out, F) through the tensor-core window GEMM."""
    g_hi, g_lo = _planes(gz)
    B, L_out, F = g_hi.shape
    check(_lib.load().bpb_input_conv_dgrad_tc(B, L_out + W - 1, L_out, W, F, _p(g_hi), _p(g_lo),
                                              _p(w_op), _p(dx), _stream()),
          "bpb_input_conv_dgrad_tc")


def heads_fwd(act, pw, pb, cw, cb, logits, gap, logcounts):
    lib = _lib.load()
    a_hi, a_lo = _planes(act)
    B, L_in, F = a_hi.shape
    W, Fw, T = pw.shape
    H = cw.shape[1]
    check(lib.bpb_heads_fwd(B, L_in, L_in - W + 1, W, Fw, F, T, H, _p(a_hi), _p(a_lo), _p(pw),
                            _p(pb), _p(cw), _p(cb), _p(logits), _p(gap), _p(logcounts),
                            _stream()), "bpb_heads_fwd")


def heads_fwd_tc_supported(W, F, T, H, planes):
    return _lib.load().bpb_heads_fwd_tc_supported(W, F, T, H, planes) == 0


def heads_fwd_groups(W, F, T, H, planes):
    """Channel groups bpb_heads_fwd_tc needs (1: weights resident; 0: unsupported)."""
    return _lib.load().bpb_heads_fwd_groups(W, F, T, H, planes)


def heads_fwd_wop(W, F, T, H, planes, device):
    n = _lib.load().bpb_heads_fwd_wop_elems(W, F, T, H, planes)
    return torch.zeros((n,), dtype=torch.bfloat16, device=device)


def prep_heads_fwd_weights(pw, cw, F, w_op, planes):
    W, Fw, T = pw.shape
    H = cw.shape[1]
    check(_lib.load().bpb_prep_heads_fwd_weights(W, Fw, F, T, H, _p(pw), _p(cw), _p(w_op), planes,
                                                 _stream()), "bpb_prep_heads_fwd_weights")


def heads_fwd_ws(B, L_in, H, device):
    n = _lib.load().bpb_heads_fwd_ws_bytes(B, L_in, H)
    return torch.zeros((n // 4 + 1,), dtype=torch.float32, device=device)


def heads_fwd_tc(act, w_op, W, T, H, pb, cb, logits, logcounts, ws):
    a_hi, a_lo = _planes(act)
    B, L_in, F = a_hi.shape
    check(_lib.load().bpb_heads_fwd_tc(B, L_in, W, F, T, H, _p(a_hi), _p(a_lo), _p(w_op), _p(pb),
                                       _p(cb), _p(logits), _p(logcounts), _p(ws),
                                       ws.numel() * ws.element_size(), _stream()),
          "bpb_heads_fwd_tc")


def heads_wgrad_ws_bytes(B, L_in, W, F, Cp):
    n = _lib.load().bpb_heads_wgrad_ws_bytes(B, L_in, W, F, Cp)
    if n < 0:
        check(1, "bpb_heads_wgrad_ws_bytes")
    return n


def heads_wgrad(act, d8, W, Fw, T, H, Cp, planes, dpw, dcw, ws):
    a_hi, a_lo = _planes(act)
    B, L_in, F = a_hi.shape
    d8_lo = (d8.data_ptr() + (d8.numel() // 2) * 2) if planes == 2 else None
    check(_lib.load().bpb_heads_wgrad_tc(B, L_in, W, Fw, F, T, H, Cp, _p(a_hi), _p(a_lo), _p(d8),
                                         d8_lo, _p(dpw), _p(dcw), _p(ws),
                                         ws.numel() * ws.element_size(), _stream()),
          "bpb_heads_wgrad_tc")


def heads_cp(T, H):
    """Channel padding of the packed head gradient d8 (multiple of 8 holding T + H values)."""
    return (T + H + 7) // 8 * 8


def heads_d8(B, L_in, W, Cp, planes, device):
    n = _lib.load().bpb_heads_d8_elems(B, L_in, W, Cp, planes)
    return torch.zeros((n,), dtype=torch.bfloat16, device=device)


def heads_dgrad_wop(W, F, Cp, planes, device):
    n = _lib.load().bpb_heads_dgrad_wop_elems(W, F, Cp, planes)
    return torch.zeros((n,), dtype=torch.bfloat16, device=device)


def prep_heads_dgrad_weights(pw, cw, F, Cp, w_op, planes):
    W, Fw, T = pw.shape
    H = cw.shape[1]
    check(_lib.load().bpb_prep_heads_dgrad_weights(W, Fw, F, T, H, Cp, _p(pw), _p(cw), _p(w_op),
                                                   planes, _stream()),
          "bpb_prep_heads_dgrad_weights")


def heads_pack_grad(dlogits, dlogcounts, L_in, W, Cp, planes, d8, dpb=None, dcb=None):
    B, L_out, T = dlogits.shape
    H = dlogcounts.shape[1]
    check(_lib.load().bpb_heads_pack_grad(B, L_in, L_out, W, T, H, Cp, planes, _p(dlogits),
                                          _p(dlogcounts), _p(d8), _p(dpb), _p(dcb), _stream()),
          "bpb_heads_pack_grad")


def heads_dgrad(d8, w_op, B, L_in, W, F, Cp, planes, g=None, gz=None, mask_in=None, mult=None):
    a = HeadsDgradArgs()
    a.B, a.L_in, a.L_out, a.W, a.F, a.Cp = B, L_in, L_in - W + 1, W, F, Cp
    a.d8_hi = _p(d8)
    a.d8_lo = (d8.data_ptr() + (d8.numel() // 2) * 2) if planes == 2 else None  # second half
    a.w_op = _p(w_op)
    g_hi, g_lo = _planes(g)
    z_hi, z_lo = _planes(gz)
    m_hi, m_lo = _planes(mult)
    a.g_hi, a.g_lo, a.gz_hi, a.gz_lo = _p(g_hi), _p(g_lo), _p(z_hi), _p(z_lo)
    a.mask_in, a.mult_hi, a.mult_lo = _p(mask_in), _p(m_hi), _p(m_lo)
    check(_lib.load().bpb_heads_dgrad_tc(C.byref(a), _stream()), "bpb_heads_dgrad_tc")


def loss_fwd_bwd(task_off, logits, logcounts, counts, profile_weight, lam, losses, dlogits=None,
                 dlogcounts=None, logcounts_true=None):
    lib = _lib.load()
    B, L, T = logits.shape
    H = logcounts.shape[1]
    check(lib.bpb_loss_fwd_bwd(B, L, T, H, _p(task_off), _p(logits), _p(logcounts), _p(counts),
                               _p(logcounts_true), _p(profile_weight), _p(lam), _p(losses),
                               _p(dlogits), _p(dlogcounts), _stream()), "bpb_loss_fwd_bwd")


def loss_pack_ws(B, T, H, device):
    """Zero-filled scratch of bpb_loss_fwd_bwd_pack (one per stream; the calls keep it consistent)."""
    n = _lib.load().bpb_loss_pack_ws_elems(B, T, H)
    return torch.zeros((n,), dtype=torch.float32, device=device)


def loss_fwd_bwd_pack(task_off, host_task_off, logits, logcounts, counts, profile_weight, lam, losses,
                      L_in, W, Cp, planes, d8, ws, dlogits=None, dlogcounts=None, dpb=None, dcb=None,
                      logcounts_true=None, counts_partials=None, cb=None):
    """K5 + bpb_heads_pack_grad in one launch; ``d8`` must have been zero-filled once
    (``heads_d8``), ``host_task_off`` is a sequence of ints."""
    lib = _lib.load()
    B, L, T = logits.shape
    H = logcounts.shape[1]
    hto = (C.c_int * (H + 1))(*[int(v) for v in host_task_off])
    check(lib.bpb_loss_fwd_bwd_pack(B, L, T, H, _p(task_off), _p(logits), _p(logcounts), _p(counts),
                                    _p(logcounts_true), _p(profile_weight), _p(lam), _p(losses),
                                    _p(dlogits), _p(dlogcounts), L_in, W, Cp, planes, _p(d8),
                                    _p(dpb), _p(dcb), _p(ws), hto, _p(counts_partials),
                                    (L_in + 127) // 128, _p(cb), _stream()),
          "bpb_loss_fwd_bwd_pack")


def adam_step(param, grad, m, v, step_and_lr, beta1=0.9, beta2=0.999, eps=1e-7, grad_scale=1.0):
    lib = _lib.load()
    check(lib.bpb_adam_step(param.numel(), _p(param), _p(grad), _p(m), _p(v), _p(step_and_lr),
                            beta1, beta2, eps, grad_scale, _stream()), "bpb_adam_step")


def ushuffle_ohe(x, kmerSize, numShuffles, seed=355687):
    """(L, A) one-hot numpy array (A <= 8) -> (numShuffles, L, A) uint8 k-let preserving shuffles
    (host; ushuffle.shuffleOHE of the reference, bit-identical for the same seed)."""
    import numpy as np
    x = np.ascontiguousarray(x, dtype=np.uint8)
    L, A = x.shape
    out = np.empty((numShuffles, L, A), dtype=np.uint8)
    check(_lib.load().bpb_ushuffle_ohe(x.ctypes.data, out.ctypes.data, A, L, int(kmerSize),
                                       int(numShuffles), int(seed) & 0xFFFFFFFF),
          "bpb_ushuffle_ohe")
    return out


def adam_prep_step(param, grad, m, v, state, table, beta1=0.9, beta2=0.999, eps=1e-7,
                   grad_scale=1.0):
    """Adam + refresh of every bf16 operand layout in one launch; ``state`` is the device vector
    {step, lr, ticket} (the kernel advances step), ``table`` a ``_lib.PrepTable``."""
    lib = _lib.load()
    check(lib.bpb_adam_prep_step(param.numel(), _p(param), _p(grad), _p(m), _p(v), _p(state),
                                 beta1, beta2, eps, grad_scale, C.byref(table), _stream()),
          "bpb_adam_prep_step")


def adam_prep_step_peer(param, m, v, state, table, peers, beta1=0.9, beta2=0.999, eps=1e-7):
    """``adam_prep_step`` with the gradient all-reduce fused in: the gradient is the sum over the
    ranks' buckets read from NVLink peer memory (``peers`` = ``_lib.PeerArgs``), scaled 1/world."""
    check(_lib.load().bpb_adam_prep_step_peer(param.numel(), _p(param), _p(m), _p(v), _p(state),
                                              beta1, beta2, eps, C.byref(table), C.byref(peers),
                                              _stream()), "bpb_adam_prep_step_peer")


def peer_wait(peers):
    """Enqueue the wait for every peer to have finished reading this rank's gradient bucket."""
    check(_lib.load().bpb_peer_wait(C.byref(peers), _stream()), "bpb_peer_wait")


def prep_conv3_weights(w, F, fwd, bwd, precise):
    """w: [3, Fw, Fw] or [layers, 3, Fw, Fw] fp32 -> fwd/bwd [layers, planes, 3, F, F] bf16."""
    lib = _lib.load()
    Fw = w.shape[-1]
    n_layers = w.shape[0] if w.dim() == 4 else 1
    check(lib.bpb_prep_conv3_weights(n_layers, Fw, F, _p(w), _p(fwd), _p(bwd),
                                     1 if precise else 0, _stream()), "bpb_prep_conv3_weights")


def onehot_encode(seq_bytes, allowN=False):
    """uint8 tensor of ASCII bases (any shape, on the device) -> (..., 4) uint8 one-hot
    (utils.oneHotEncode); raises like the reference when allowN is False and a byte is not ACGT."""
    seq_bytes = seq_bytes.contiguous()
    out = torch.empty(seq_bytes.shape + (4,), dtype=torch.uint8, device=seq_bytes.device)
    cnt = torch.zeros((1,), dtype=torch.int64, device=seq_bytes.device)
    check(_lib.load().bpb_onehot_encode(seq_bytes.numel(), _p(seq_bytes), _p(out), _p(cnt), _stream()),
          "bpb_onehot_encode")
    if not allowN and int(cnt.item()) != 0:
        raise AssertionError("Sequence contains unrecognized nucleotides. Maybe your sequence "
                             "contains 'N'?")
    return out


def logits_to_profile(task_off, logits, logcounts, profile=None):
    """(B, L, T) logits, (B, H) log-counts -> (B, L, T) predicted profiles (utils.logitsToProfile
    applied to every head of every sequence)."""
    B, L, T = logits.shape
    H = logcounts.shape[1]
    if profile is None:
        profile = torch.empty_like(logits)
    check(_lib.load().bpb_logits_to_profile(B, L, T, H, _p(task_off), _p(logits.contiguous()),
                                            _p(logcounts.contiguous()), _p(profile), _stream()),
          "bpb_logits_to_profile")
    return profile


def slide_u8(src, dst, row_idx, col_idx):
    lib = _lib.load()
    rows, src_cols, depth = src.shape
    check(lib.bpb_slide_u8(_p(src), _p(dst), rows, src_cols, dst.shape[1], depth, _p(row_idx),
                           _p(col_idx), _stream()), "bpb_slide_u8")


def slide_f32(src, dst, row_idx, col_idx):
    lib = _lib.load()
    rows, src_cols, depth = src.shape
    check(lib.bpb_slide_f32(_p(src), _p(dst), rows, src_cols, dst.shape[1], depth, _p(row_idx),
                            _p(col_idx), _stream()), "bpb_slide_f32")


def log_row_sums(vals, out):
    lib = _lib.load()
    rows = vals.shape[0]
    check(lib.bpb_log_row_sums(_p(vals), _p(out), rows, vals.numel() // rows, _stream()),
          "bpb_log_row_sums")


def transform_fwd(u, y, act, coeffs):
    lib = _lib.load()
    check(lib.bpb_transform_fwd(u.numel(), _p(u), _p(y), act.numel(), _p(act), _p(coeffs),
                                _stream()), "bpb_transform_fwd")


def transform_bwd(u, dy, act, coeffs, dcoeffs):
    lib = _lib.load()
    check(lib.bpb_transform_bwd(u.numel(), _p(u), _p(dy), act.numel(), _p(act), _p(coeffs),
                                _p(dcoeffs), _stream()), "bpb_transform_bwd")


def add_f32(a, b, out):
    check(_lib.load().bpb_add_f32(a.numel(), _p(a), _p(b), _p(out), _stream()), "bpb_add_f32")


def mul_f32(a, b, out):
    check(_lib.load().bpb_mul_f32(a.numel(), _p(a), _p(b), _p(out), _stream()), "bpb_mul_f32")


def counts_lse(use_bias, bias_lc, resid_lc, out, dscale=None):
    B, H = resid_lc.shape
    check(_lib.load().bpb_counts_lse(B, H, _p(use_bias), _p(bias_lc), _p(resid_lc), _p(out),
                                     _p(dscale), _stream()), "bpb_counts_lse")


def combine_bias(task_off, bias_logits, bias_lc, p_act, c_act, coeffs, n_p, n_c, use_bias,
                 resid_logits, resid_lc, out_logits, out_lc, dscale=None, bias_logits_t=None,
                 bias_lc_t=None):
    """bpb_combine_bias: transformation of the bias outputs + logit add + CountsLogSumExp."""
    B, L, T = resid_logits.shape
    H = resid_lc.shape[1]
    check(_lib.load().bpb_combine_bias(B, L, T, H, _p(task_off), _p(bias_logits), _p(bias_lc),
                                       n_p, _p(p_act), n_c, _p(c_act), _p(coeffs), _p(use_bias),
                                       _p(resid_logits), _p(resid_lc), _p(out_logits), _p(out_lc),
                                       _p(dscale), _p(bias_logits_t), _p(bias_lc_t), _stream()),
          "bpb_combine_bias")


def combined_shap_seed(n, task_off, bl_x, bl_r, bc_x, bc_r, rc_x, rc_r, p_act, c_act, coeffs, n_p,
                       n_c, use_bias, dl_c, dc_c, dl_r=None, dc_r=None, dl_b=None, dc_b=None):
    """bpb_combined_shap_seed: rescale-rule gradients through transformation + add + LSE."""
    R, L, T = bl_r.shape
    H = bc_r.shape[1]
    check(_lib.load().bpb_combined_shap_seed(
        R, n, L, T, H, _p(task_off), _p(bl_x), _p(bl_r), _p(bc_x), _p(bc_r), _p(rc_x), _p(rc_r),
        n_p, _p(p_act), n_c, _p(c_act), _p(coeffs), _p(use_bias), _p(dl_c), _p(dc_c), _p(dl_r),
        _p(dc_r), _p(dl_b), _p(dc_b), _stream()), "bpb_combined_shap_seed")


def shuffle_rows(x, perm, refs):
    Q, L, _ = x.shape
    n = perm.shape[0]
    check(_lib.load().bpb_shuffle_rows(Q, n, L, _p(x), _p(perm), _p(refs), _stream()),
          "bpb_shuffle_rows")


def flat_profile_seed(sel, logits_x, logits_r, dlogits, vx=None, vr=None):
    Q, L, T = logits_x.shape
    n = logits_r.shape[0] // Q
    check(_lib.load().bpb_flat_profile_seed(Q, n, L, T, _p(sel), _p(logits_x), _p(logits_r),
                                            _p(dlogits), _p(vx), _p(vr), _stream()),
          "bpb_flat_profile_seed")


def shap_combine(x, refs, mult, phi, hypothetical=False):
    Q, L, _ = x.shape
    n = refs.shape[0] // Q
    check(_lib.load().bpb_shap_combine(Q, n, L, _p(x), _p(refs), _p(mult),
                                       1 if hypothetical else 0, _p(phi), _stream()),
          "bpb_shap_combine")
