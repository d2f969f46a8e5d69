"""Diagnostic (not collected): where do repeated launches of one tower layer differ?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bpreveal_b200 import ops  # noqa: E402
from tests.test_gpu_baseline_cfgs import prep_w, split  # noqa: E402

dev = torch.device("cuda:0")
B, Fn = 64, 64
g = torch.Generator().manual_seed(9)
cases = [(2942, 64)] if len(sys.argv) < 2 else [(int(a.split(",")[0]), int(a.split(",")[1])) for a in sys.argv[1:]]
for (L, d) in cases:
    L_out = L - 2 * d
    xf = torch.randn((B, L, Fn), generator=g).to(dev)
    x = split(xf, False)
    w = (torch.randn((3, Fn, Fn), generator=g) / 14).to(dev)
    fwd, bwd = prep_w(w, Fn, False, dev)
    out = ops.new_planes(B, L_out, Fn, False, dev)
    mask = torch.zeros((B, L_out, 1), dtype=torch.int64, device=dev)
    bias = torch.zeros((Fn,), device=dev)

    def f_fwd():
        ops.conv3(x, fwd, (0, d, 2 * d), 0, L_out, bias=bias, resid=x, resid_off=d, out=out,
                  mask_out=mask)

    f_fwd()
    torch.cuda.synchronize()
    first = out[0].clone()
    xe = x[0].float()
    we = w.to(torch.bfloat16).float()
    z = torch.nn.functional.conv1d(xe.transpose(1, 2), we.permute(2, 1, 0), dilation=d).transpose(1, 2)
    resid = xe[:, d:L - d, :]
    ref = torch.relu(z) + resid
    print("first vs ref max err", float((first.float() - ref).abs().max()), flush=True)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(10):
            f_fwd()
    bad = 0
    for rep in range(150):
        gr.replay()
        torch.cuda.synchronize()
        now = out[0]
        if not torch.equal(first, now):
            bad += 1
            idx = (first != now).nonzero()
            print(f"d={d} rep {rep}: {idx.shape[0]} elements differ", flush=True)
            rows = {}
            for b_, t_, c_ in idx.tolist():
                rows.setdefault((b_, t_), []).append(c_)
            for (b_, t_), cs in sorted(rows.items())[:40]:
                c0 = cs[0]
                print(f"   b={b_} t={t_} (tile {t_ // 128} row {t_ % 128}) ch {cs[0]}..{cs[-1]} n={len(cs)} "
                      f"first={float(first[b_, t_, c0]):.4f} now={float(now[b_, t_, c0]):.4f} "
                      f"ref={float(ref[b_, t_, c0]):.4f} relu(z)={float(torch.relu(z[b_, t_, c0])):.4f} "
                      f"resid={float(resid[b_, t_, c0]):.4f}", flush=True)
            # where does the wrong residual come from?  (rows whose relu(z) is exactly 0: out = resid)
            xs = x[0]
            shown = 0
            for (b_, t_), cs in sorted(rows.items()):
                c0, c1 = cs[0], cs[-1] + 1
                bad_t = first if not torch.equal(first[b_, t_, c0:c1], ref[b_, t_, c0:c1].to(torch.bfloat16)) else now
                if float(torch.relu(z[b_, t_, c0:c1]).abs().max()) != 0.0:
                    continue
                pat = bad_t[b_, t_, c0:c1]
                hit = (xs[:, :, c0:c1] == pat).all(dim=2).nonzero()
                print(f"   wrong resid of b={b_} t={t_} ch {c0}..{c1 - 1} (should be x[{b_},{t_ + d}]) "
                      f"equals x at {hit.tolist()[:4]}", flush=True)
                shown += 1
                if shown >= 6:
                    break
    print(f"d={d}: {bad} differing replays of 150", flush=True)
