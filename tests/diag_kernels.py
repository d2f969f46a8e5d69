"""Standalone kernel diagnostics (not collected by pytest): each group compares one native kernel
against a plain fp32 torch computation on the same (bf16-rounded) inputs and prints max errors.
Usage: python tests/diag_kernels.py <group> ...   (run each group under `timeout`)."""
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bpreveal_b200 import ops  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")


def conv_cl(x, w, dilation=1):
    """x (B,L,C) fp32, w (k,Cin,Cout) -> (B, L-d(k-1), Cout), cross-correlation."""
    return F.conv1d(x.transpose(1, 2), w.permute(2, 1, 0), dilation=dilation).transpose(1, 2)


def split(v, precise):
    hi = v.to(torch.bfloat16)
    lo = (v - hi.float()).to(torch.bfloat16) if precise else None
    return (hi.contiguous(), lo.contiguous() if lo is not None else None)


def val(ps):
    return ops.planes_to_float(ps)


def report(name, got, ref, tol):
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-30
    ok = err <= tol * scale
    print(f"  {name}: max_abs_err={err:.3e} ref_max={scale:.3e} rel={err / scale:.3e} "
          f"{'OK' if ok else 'FAIL'}", flush=True)
    return ok


def prep_w(w, Fp, precise):
    n = 2 if precise else 1
    fwd = torch.empty((n, 3, Fp, Fp), dtype=torch.bfloat16, device=dev)
    bwd = torch.empty((n, 3, Fp, Fp), dtype=torch.bfloat16, device=dev)
    ops.prep_conv3_weights(w.contiguous(), Fp, fwd, bwd, precise)
    return fwd, bwd


def w_eff(w, precise):
    """The weight values the kernel actually multiplies with."""
    hi = w.to(torch.bfloat16).float()
    if precise:
        return hi + (w - hi).to(torch.bfloat16).float()
    return hi


def group_conv_fwd(cases=None):
    ok = True
    cases = cases or [(3, 300, 64, 2, False), (2, 1100, 64, 64, False), (2, 700, 128, 8, False),
                      (1, 600, 256, 4, False), (2, 300, 64, 4, True), (1, 700, 128, 16, True),
                      (2, 52, 64, 2, False), (1, 500, 192, 2, False), (2, 1200, 64, 256, False),
                      (2, 900, 64, 128, True), (3, 3066, 64, 2, False), (2, 700, 128, 128, False)]
    for (B, L, Fn, d, precise) in cases:
        print(f"conv fwd B={B} L={L} F={Fn} d={d} precise={precise}", flush=True)
        g = torch.Generator(device="cpu").manual_seed(1)
        x = torch.randn((B, L, Fn), generator=g).to(dev)
        w = (torch.randn((3, Fn, Fn), generator=g) / (3 * Fn) ** 0.5).to(dev)
        bias = torch.randn((Fn,), generator=g).to(dev) * 0.1
        xin = split(x, precise)
        fwd, _ = prep_w(w, Fn, precise)
        L_out = L - 2 * d
        out = ops.new_planes(B, L_out, Fn, precise, dev)
        mask = torch.zeros((B, L_out, Fn // 64), dtype=torch.int64, device=dev)
        z = torch.empty((B, L_out, Fn), dtype=torch.float32, device=dev)
        ops.conv3(xin, fwd, (0, d, 2 * d), 0, L_out, bias=bias, resid=xin, resid_off=d, out=out,
                  mask_out=mask, z_out=z)
        torch.cuda.synchronize()
        xe = val(xin)
        zr = conv_cl(xe, w_eff(w, precise), d) + bias
        ref = torch.relu(zr) + xe[:, d:L - d, :]
        tol = 2e-5 if precise else 6e-3
        ok &= report("z", z, zr, 1e-4 if precise else 1e-5 * 10)
        ok &= report("out", val(out), ref, tol)
        bits = torch.stack([(mask[..., c // 64] >> (c % 64)) & 1 for c in range(Fn)], dim=-1)
        mism = (bits.bool() != (z > 0)).sum().item()
        print(f"  mask mismatches vs own z: {mism}", flush=True)
        ok &= mism == 0
    return ok


def group_conv_bwd():
    ok = True
    for (B, L_out_f, Fn, d, precise, use_mult) in [(3, 300, 64, 2, False, False),
                                                   (2, 900, 64, 128, False, True),
                                                   (2, 400, 128, 8, True, False),
                                                   (1, 300, 256, 4, False, False),
                                                   (2, 300, 64, 4, True, True),
                                                   (3, 1000, 64, 32, False, False),
                                                   (2, 500, 64, 64, False, False),
                                                   (2, 1022, 64, 512, False, False),
                                                   (2, 300, 128, 16, False, True)]:
        # gz has length L_out_f (forward output), result has length L_out_f + 2d
        print(f"conv bwd B={B} Lg={L_out_f} F={Fn} d={d} precise={precise} mult={use_mult}",
              flush=True)
        g = torch.Generator(device="cpu").manual_seed(2)
        L_in = L_out_f + 2 * d
        gz = torch.randn((B, L_out_f, Fn), generator=g).to(dev)
        gp = torch.randn((B, L_out_f, Fn), generator=g).to(dev)
        w = (torch.randn((3, Fn, Fn), generator=g) / (3 * Fn) ** 0.5).to(dev)
        gzp, gpp = split(gz, precise), split(gp, precise)
        _, bwd = prep_w(w, Fn, precise)
        out = ops.new_planes(B, L_in, Fn, precise, dev)
        out2 = ops.new_planes(B, L_in, Fn, precise, dev)
        maskb = (torch.rand((B, L_in, Fn), generator=g) > 0.5).to(dev)
        mask = torch.zeros((B, L_in, Fn // 64), dtype=torch.int64, device=dev)
        for c in range(Fn):
            mask[..., c // 64] |= maskb[..., c].long() << (c % 64)
        multv = torch.rand((B, L_in, Fn), generator=g).to(dev)
        mult = split(multv, precise)
        ops.conv3(gzp, bwd, (0, -d, -2 * d), 1, L_in, resid=gpp, resid_off=-d, out=out, out2=out2,
                  mask_in=None if use_mult else mask, mult=mult if use_mult else None)
        torch.cuda.synchronize()
        # reference via autograd of the forward conv
        a = torch.zeros((B, L_in, Fn), device=dev, requires_grad=True)
        zf = conv_cl(a, w_eff(w, precise), d)
        (zf * val(gzp)).sum().backward()
        ref = a.grad.clone()
        ref[:, d:L_in - d, :] += val(gpp)
        tol = 2e-5 if precise else 6e-3
        ok &= report("g'", val(out), ref, tol)
        ref2 = ref * (val(mult) if use_mult else maskb.float())
        ok &= report("gz", val(out2), ref2, tol)
    return ok


def group_wgrad():
    ok = True
    for (B, L_in, Fn, d, precise) in [(3, 300, 64, 2, False), (4, 1100, 64, 128, False),
                                      (2, 500, 128, 8, False), (2, 400, 64, 4, True),
                                      (1, 400, 256, 4, False), (64, 1100, 64, 32, False)]:
        print(f"wgrad B={B} L_in={L_in} F={Fn} d={d} precise={precise}", flush=True)
        g = torch.Generator(device="cpu").manual_seed(3)
        L_out = L_in - 2 * d
        a = split(torch.randn((B, L_in, Fn), generator=g).to(dev), precise)
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
        ok &= report("dw", dw, ref, tol)
        ok &= report("db", db, gv.sum(dim=(0, 1)), tol)
    return ok


def group_edges():
    ok = True
    g = torch.Generator(device="cpu").manual_seed(4)
    for (B, L_in, W, Fw, Fp, precise) in [(3, 400, 25, 64, 64, False), (2, 300, 7, 16, 64, True),
                                          (2, 300, 3, 8, 64, False)]:
        print(f"input conv B={B} L={L_in} W={W} Fw={Fw} F={Fp} precise={precise}", flush=True)
        idx = torch.randint(0, 4, (B, L_in), generator=g)
        x = F.one_hot(idx, 4).to(torch.uint8).to(dev)
        w = (torch.randn((W, 4, Fw), generator=g) * 0.3).to(dev)
        bias = (torch.randn((Fw,), generator=g) * 0.1).to(dev)
        L_out = L_in - W + 1
        out = ops.new_planes(B, L_out, Fp, precise, dev)
        mask = torch.zeros((B, L_out, Fp // 64), dtype=torch.int64, device=dev)
        z = torch.empty((B, L_out, Fp), dtype=torch.float32, device=dev)
        planes = 2 if precise else 1
        x8 = ops.new_x8(B, L_in, dev)
        ops.onehot_expand(x, x8)
        assert torch.equal(x8[..., :4].float(), x.float()) and x8[..., 4:].abs().max() == 0
        w_op = ops.input_conv_wop(W, Fp, planes, dev)
        ops.prep_input_conv_weights(w, Fp, w_op, planes)
        ops.input_conv_fwd(x8, w_op, planes, W, Fw, bias, Fp, out, mask_out=mask, z_out=z)
        torch.cuda.synchronize()
        zr = conv_cl(x.float(), w_eff(w, precise)) + bias
        ok &= report("z", z[..., :Fw], zr, 2e-5)
        ok &= report("z vs fp32 weights", z[..., :Fw], conv_cl(x.float(), w) + bias,
                     2e-5 if precise else 1e-2)
        ok &= report("out", val(out)[..., :Fw], torch.relu(zr), 2e-5 if precise else 5e-3)
        if Fp > Fw:
            print("  pad channels max:", val(out)[..., Fw:].abs().max().item())
        bits = torch.stack([(mask[..., c // 64] >> (c % 64)) & 1 for c in range(Fw)], dim=-1)
        mism = (bits.bool() != (z[..., :Fw] > 0)).sum().item()
        print(f"  mask mismatches vs own z: {mism}")
        ok &= mism == 0
        # wgrad
        gz = split(torch.randn((B, L_out, Fp), generator=g).to(dev), precise)
        dw = torch.empty((W, 4, Fw), dtype=torch.float32, device=dev)
        db = torch.empty((Fw,), dtype=torch.float32, device=dev)
        ws = torch.empty((ops.input_conv_wgrad_ws_bytes(B, L_out, W, Fp) // 4,),
                         dtype=torch.float32, device=dev)
        ops.input_conv_wgrad(x8, gz, W, Fw, dw, db, ws)
        gv = val(gz)[..., :Fw]
        xf = x.float()
        ref = torch.stack([torch.einsum("btc,btf->cf", xf[:, j:j + L_out, :], gv)
                           for j in range(W)])
        ok &= report("dw1", dw, ref, 1e-4)
        ok &= report("db1", db, gv.sum(dim=(0, 1)), 1e-4)
        assert dw.abs().max() > 0
        # dgrad
        dx = torch.empty((B, L_in, 4), dtype=torch.float32, device=dev)
        ops.input_conv_dgrad(gz, w, L_in, dx)
        xa = torch.zeros((B, L_in, 4), device=dev, requires_grad=True)
        (conv_cl(xa, w) * gv).sum().backward()
        ok &= report("dx", dx, xa.grad, 1e-4)

    for (B, L_in, W, Fw, Fp, tasks, precise) in [(3, 400, 23, 64, 64, [2], False),
                                                 (2, 300, 75, 64, 64, [2, 2, 2, 2], False),
                                                 (2, 200, 7, 16, 64, [1, 3], True)]:
        print(f"heads B={B} L={L_in} W={W} Fw={Fw} tasks={tasks} precise={precise}", flush=True)
        T, H = sum(tasks), len(tasks)
        L_out = L_in - W + 1
        av = torch.randn((B, L_in, Fp), generator=g).to(dev)
        av[..., Fw:] = 0
        act = split(av, precise)
        pw = (torch.randn((W, Fw, T), generator=g) * 0.1).to(dev)
        pb = (torch.randn((T,), generator=g) * 0.1).to(dev)
        cw = (torch.randn((Fw, H), generator=g) * 0.1).to(dev)
        cb = (torch.randn((H,), generator=g) * 0.1).to(dev)
        logits = torch.empty((B, L_out, T), dtype=torch.float32, device=dev)
        gap = torch.empty((B, Fp), dtype=torch.float32, device=dev)
        lc = torch.empty((B, H), dtype=torch.float32, device=dev)
        ops.heads_fwd(act, pw, pb, cw, cb, logits, gap, lc)
        ae = val(act)[..., :Fw]
        lr = conv_cl(ae, pw) + pb
        ok &= report("logits", logits, lr, 1e-4)
        # the same through the tensor-core kernel
        planes_h = 2 if precise else 1
        T_, H_ = sum(tasks), len(tasks)
        if ops.heads_fwd_tc_supported(W, Fp, T_, H_, planes_h):
            w_hf = ops.heads_fwd_wop(W, Fp, T_, H_, planes_h, dev)
            ops.prep_heads_fwd_weights(pw, cw, Fp, w_hf, planes_h)
            logits2 = torch.empty_like(logits)
            lc2 = torch.empty_like(lc)
            ops.heads_fwd_tc(act, w_hf, W, T_, H_, pb, cb, logits2, lc2,
                             ops.heads_fwd_ws(B, L_in, H_, dev))
            ok &= report("logits (tc)", logits2, lr, 2e-5 if precise else 4e-3)
            ok &= report("logcounts (tc)", lc2, ae.mean(dim=1) @ cw + cb, 2e-5 if precise else 4e-3)
        else:
            print("  heads_fwd_tc not supported for this geometry", flush=True)
        gr = ae.mean(dim=1)
        ok &= report("gap", gap[:, :Fw], gr, 1e-4)
        ok &= report("logcounts", lc, gr @ cw + cb, 1e-4)
        # losses
        counts = torch.poisson(torch.rand((B, L_out, T), generator=g) * 2).to(dev)
        counts[:, 0, :] += 1
        toff = torch.tensor([0] + list(torch.tensor(tasks).cumsum(0)), dtype=torch.int32,
                            device=dev)
        pwt = torch.rand((H,), generator=g).to(dev) + 0.5
        lam = torch.rand((H,), generator=g).to(dev) + 0.5
        losses = torch.empty((2 * H,), dtype=torch.float32, device=dev)
        dlog = torch.empty_like(logits)
        dlc = torch.empty_like(lc)
        ops.loss_fwd_bwd(toff, logits, lc, counts, pwt, lam, losses, dlog, dlc)
        lg = logits.clone().requires_grad_(True)
        lcr = lc.clone().requires_grad_(True)
        tot = 0
        ref_losses = []
        for h in range(H):
            k0, k1 = int(toff[h]), int(toff[h + 1])
            kk = counts[:, :, k0:k1].reshape(B, -1)
            ll = lg[:, :, k0:k1].reshape(B, -1)
            N = kk.sum(1)
            logp = torch.lgamma(N + 1) - torch.lgamma(kk + 1).sum(1) + \
                (kk * torch.log_softmax(ll, 1)).sum(1)
            mn = -logp.sum() / B
            ms = ((lcr[:, h] - torch.log(N)) ** 2).mean() * lam[h]
            ref_losses.append((mn, ms))
            tot = tot + pwt[h] * mn + ms
        tot.backward()
        ok &= report("mnll", losses[:H], torch.stack([r[0] for r in ref_losses]).detach(), 1e-5)
        ok &= report("wmse", losses[H:], torch.stack([r[1] for r in ref_losses]).detach(), 1e-5)
        ok &= report("dlogits", dlog, lg.grad, 1e-4)
        ok &= report("dlogcounts", dlc, lcr.grad, 1e-4)
        # heads backward
        gpl = ops.new_planes(B, L_in, Fp, precise, dev)
        gzp = ops.new_planes(B, L_in, Fp, precise, dev)
        maskb = (torch.rand((B, L_in, Fp), generator=g) > 0.5).to(dev)
        mask = torch.zeros((B, L_in, Fp // 64), dtype=torch.int64, device=dev)
        for c in range(Fp):
            mask[..., c // 64] |= maskb[..., c].long() << (c % 64)
        dpw, dpb = torch.empty_like(pw), torch.empty_like(pb)
        dcw, dcb = torch.empty_like(cw), torch.empty_like(cb)
        planes = 2 if precise else 1
        Cp = ops.heads_cp(T, H)
        d8 = ops.heads_d8(B, L_in, W, Cp, planes, dev)
        ws = torch.empty((ops.heads_wgrad_ws_bytes(B, L_in, W, Fp, Cp) // 4,), dtype=torch.float32,
                         device=dev)
        w_hd = ops.heads_dgrad_wop(W, Fp, Cp, planes, dev)
        ops.prep_heads_dgrad_weights(pw, cw, Fp, Cp, w_hd, planes)
        ops.heads_pack_grad(dlog, dlc, L_in, W, Cp, planes, d8, dpb=dpb, dcb=dcb)
        ops.heads_wgrad(act, d8, W, Fw, T, H, Cp, planes, dpw, dcw, ws)
        ops.heads_dgrad(d8, w_hd, B, L_in, W, Fp, Cp, planes, g=gpl, gz=gzp, mask_in=mask)
        aa = ae.clone().requires_grad_(True)
        pwr, pbr = pw.clone().requires_grad_(True), pb.clone().requires_grad_(True)
        cwr, cbr = cw.clone().requires_grad_(True), cb.clone().requires_grad_(True)
        l2 = conv_cl(aa, pwr) + pbr
        c2 = aa.mean(1) @ cwr + cbr
        ((l2 * dlog).sum() + (c2 * dlc).sum()).backward()
        tolg = 3e-5 if precise else 8e-3
        ok &= report("g", val(gpl)[..., :Fw], aa.grad, tolg)
        ok &= report("gz", val(gzp)[..., :Fw], aa.grad * maskb[..., :Fw].float(), tolg)
        tolw = 1e-4 if precise else 5e-3  # d8 is the loss gradient rounded to bf16 (x2 if precise)
        ok &= report("dpw", dpw, pwr.grad, tolw)
        ok &= report("dpb", dpb, pbr.grad, 1e-4)
        ok &= report("dcw", dcw, cwr.grad, tolw)
        ok &= report("dcb", dcb, cbr.grad, 1e-4)
    # adam
    n = 10007
    p = torch.randn((n,), generator=g).to(dev)
    gr = torch.randn((n,), generator=g).to(dev)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    pr, mr, vr = p.double().clone(), m.double().clone(), v.double().clone()
    sl = torch.zeros((2,), dtype=torch.float32, device=dev)
    for step in range(1, 4):
        sl[0], sl[1] = float(step), 0.004
        ops.adam_step(p, gr, m, v, sl)
        alpha = 0.004 * (1 - 0.999 ** step) ** 0.5 / (1 - 0.9 ** step)
        mr += (gr.double() - mr) * (1 - 0.9)
        vr += (gr.double() ** 2 - vr) * (1 - 0.999)
        pr -= alpha * mr / (vr.sqrt() + 1e-7)
    ok &= report("adam p", p, pr.float(), 1e-6)
    return ok


def group_bench():
    """Rough timing of the tower kernels at the C1 shape."""
    B, Fn = 64, 64
    for (L, d) in [(3066, 2), (3006, 32), (2046, 512)]:
        x = torch.randn((B, L, Fn), device=dev)
        w = torch.randn((3, Fn, Fn), device=dev) / 14
        xin = split(x, False)
        fwd, bwd = prep_w(w, Fn, False)
        L_out = L - 2 * d
        out = ops.new_planes(B, L_out, Fn, False, dev)
        mask = torch.zeros((B, L_out, 1), dtype=torch.int64, device=dev)
        bias = torch.zeros((Fn,), device=dev)
        dw = torch.empty((3, Fn, Fn), dtype=torch.float32, device=dev)
        db = torch.empty((Fn,), dtype=torch.float32, device=dev)
        ws = torch.empty((ops.wgrad3_ws_bytes(B, L_out, Fn, False) // 4,), dtype=torch.float32,
                         device=dev)
        gin = ops.new_planes(B, L, Fn, False, dev)
        gz2 = ops.new_planes(B, L, Fn, False, dev)
        mask_in = torch.zeros((B, L, 1), dtype=torch.int64, device=dev)

        def f_fwd():
            ops.conv3(xin, fwd, (0, d, 2 * d), 0, L_out, bias=bias, resid=xin, resid_off=d,
                      out=out, mask_out=mask)

        def f_bwd():
            ops.conv3(out, bwd, (0, -d, -2 * d), 1, L, resid=out, resid_off=-d, out=gin,
                      out2=gz2, mask_in=mask_in)

        def f_wg():
            ops.wgrad3(xin, out, d, dw, db, ws)

        for name, fn in (("fwd", f_fwd), ("dgrad", f_bwd), ("wgrad", f_wg)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            # device time: replay a CUDA graph of 20 launches (host-side launch cost excluded)
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(20):
                    fn()
            gr.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gr.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 20 * 1e3
            fl = 2.0 * B * L_out * 3 * Fn * Fn
            by = 2.0 * B * (L + L_out) * Fn
            print(f"L={L} d={d} {name}: {us:.1f} us  {fl / us / 1e6:.1f} TFLOP/s  "
                  f"{by / us / 1e3:.0f} GB/s(alg)", flush=True)
    return True


def group_bench512():
    """C4-like layers (F=512, 9262-bp activations): the tensor-bound regime."""
    B, Fn = 8, 512
    for (L, d) in [(9262, 2), (9000, 64), (6000, 1024)]:
        x = torch.randn((B, L, Fn), device=dev)
        w = torch.randn((3, Fn, Fn), device=dev) / 40
        xin = split(x, False)
        fwd, bwd = prep_w(w, Fn, False)
        L_out = L - 2 * d
        out = ops.new_planes(B, L_out, Fn, False, dev)
        mask = torch.zeros((B, L_out, Fn // 64), dtype=torch.int64, device=dev)
        bias = torch.zeros((Fn,), device=dev)
        dw = torch.empty((3, Fn, Fn), dtype=torch.float32, device=dev)
        db = torch.empty((Fn,), dtype=torch.float32, device=dev)
        ws = torch.empty((ops.wgrad3_ws_bytes(B, L_out, Fn, False) // 4,), dtype=torch.float32,
                         device=dev)
        gin = ops.new_planes(B, L, Fn, False, dev)
        gz2 = ops.new_planes(B, L, Fn, False, dev)
        mask_in = torch.zeros((B, L, Fn // 64), dtype=torch.int64, device=dev)

        def f_fwd():
            ops.conv3(xin, fwd, (0, d, 2 * d), 0, L_out, bias=bias, resid=xin, resid_off=d,
                      out=out, mask_out=mask)

        def f_bwd():
            ops.conv3(out, bwd, (0, -d, -2 * d), 1, L, resid=out, resid_off=-d, out=gin,
                      out2=gz2, mask_in=mask_in)

        def f_wg():
            ops.wgrad3(xin, out, d, dw, db, ws)

        for name, fn in (("fwd", f_fwd), ("dgrad", f_bwd), ("wgrad", f_wg)):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(5):
                    fn()
            gr.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gr.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 5 * 1e3
            fl = 2.0 * B * L_out * 3 * Fn * Fn
            print(f"F=512 L={L} d={d} {name}: {us:.1f} us  {fl / us / 1e6:.1f} TFLOP/s", flush=True)
    return True


def group_stress():
    """Every C1 tower layer shape, forward and backward, many launches (races show up here)."""
    B, Fn = 64, 64
    L = 3066
    for i in range(9):
        d = 2 ** (i + 1)
        L_out = L - 2 * d
        x = torch.randn((B, L, Fn), device=dev)
        w = torch.randn((3, Fn, Fn), device=dev) / 14
        xin = split(x, False)
        fwd, bwd = prep_w(w, Fn, False)
        out = ops.new_planes(B, L_out, Fn, False, dev)
        mask = torch.zeros((B, L_out, 1), dtype=torch.int64, device=dev)
        bias = torch.zeros((Fn,), device=dev)
        gin = ops.new_planes(B, L, Fn, False, dev)
        gz2 = ops.new_planes(B, L, Fn, False, dev)
        mask_in = torch.zeros((B, L, 1), dtype=torch.int64, device=dev)
        def f_fwd():
            ops.conv3(xin, fwd, (0, d, 2 * d), 0, L_out, bias=bias, resid=xin, resid_off=d,
                      out=out, mask_out=mask)

        def f_bwd():
            ops.conv3(out, bwd, (0, -d, -2 * d), 1, L, resid=out, resid_off=-d,
                      out=gin, out2=gz2, mask_in=mask_in)

        def f_bwd0():
            ops.conv3(out, bwd, (0, -d, -2 * d), 1, L, resid=out, resid_off=-d,
                      out=None, out2=gz2, mask_in=mask_in)

        for name, fn in (("fwd", f_fwd), ("bwd", f_bwd), ("bwd-noout", f_bwd0)):
            fn()
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(20):
                    fn()
            for _ in range(10):
                gr.replay()
            torch.cuda.synchronize()
            print(f"layer {i} d={d} {name} ok", flush=True)
        L = L_out
    return True


if __name__ == "__main__":
    groups = {"conv_fwd": group_conv_fwd, "conv_bwd": group_conv_bwd, "wgrad": group_wgrad,
              "edges": group_edges, "bench": group_bench, "stress": group_stress, "bench512": group_bench512}
    allok = True
    for name in sys.argv[1:]:
        print(f"=== {name}", flush=True)
        t0 = time.time()
        try:
            r = groups[name]()
        except Exception as e:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            from bpreveal_b200 import _lib
            for rec in _lib.debug_dump()[:24]:
                print("  watchdog: tag=0x%x block=%d thread=%d parity=%d tile=%d" % rec, flush=True)
            r = False
        allok &= bool(r)
        print(f"=== {name} -> {'PASS' if r else 'FAIL'} ({time.time() - t0:.1f}s)", flush=True)
    sys.exit(0 if allok else 1)
