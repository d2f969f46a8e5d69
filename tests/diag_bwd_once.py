"""One launch each of the tower's data gradient (materialised and masked-operand path) and the
forward at a given dilation, C1 shapes at batch 64: the target of `ncu --set full` captures.
Usage: diag_bwd_once.py [dilation ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bpreveal_b200 import ops  # noqa: E402
from tests.diag_kernels import prep_w, split  # noqa: E402

dev = torch.device("cuda:0")
B, Fn, L = 64, 64, 3066
for d in [int(a) for a in sys.argv[1:]] or [2, 64, 256]:
    Lg = L - 2 * d
    g = split(torch.randn((B, Lg, Fn), device=dev), False)
    x = split(torch.randn((B, L, Fn), device=dev), False)
    w = torch.randn((3, Fn, Fn), device=dev) / 14
    fwd, bwd = prep_w(w, Fn, False)
    out = ops.new_planes(B, L, Fn, False, dev)
    out2 = ops.new_planes(B, L, Fn, False, dev)
    outf = ops.new_planes(B, Lg, Fn, False, dev)
    bias = torch.zeros((Fn,), device=dev)
    mask_o = torch.zeros((B, Lg, 1), dtype=torch.int64, device=dev)
    mask_in = torch.randint(-2 ** 62, 2 ** 62, (B, L, 1), dtype=torch.int64, device=dev)
    mask_a = torch.randint(-2 ** 62, 2 ** 62, (B, Lg, 1), dtype=torch.int64, device=dev)
    for _ in range(2):
        ops.conv3(x, fwd, (0, d, 2 * d), 0, Lg, bias=bias, resid=x, resid_off=d, out=outf,
                  mask_out=mask_o)
        ops.conv3(g, bwd, (0, -d, -2 * d), 1, L, resid=g, resid_off=-d, out=out, out2=out2,
                  mask_in=mask_in)
        ops.conv3(g, bwd, (0, -d, -2 * d), 1, L, resid=g, resid_off=-d, out=out, mask_a=mask_a)
    torch.cuda.synchronize()
print("ok")
