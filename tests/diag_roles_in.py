"""Role profile (library built with `make PROFILE=1`) of the input convolution at C1 size."""
import ctypes as C
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bpreveal_b200 import _lib, ops  # noqa: E402
from tests.diag_roles import NAMES  # noqa: E402,F401

dev = torch.device("cuda:0")


def snapshot():
    buf = (C.c_uint * 2048)()
    _lib.load().bpb_debug_dump(buf, 2048)
    return buf


def report(label, b0, b1, n):
    print(label)
    for cta in range(2):
        base = 1024 + 32 * cta
        d = [(b1[base + i] - b0[base + i]) / n for i in range(32)]
        print(f"  cta{cta}: total={d[15]:.0f} prod:empty={d[1]:.0f} mma:w_full={d[3]:.0f} "
              f"mma:acc_empty={d[4]:.0f} mma:full={d[5]:.0f} epi:halo={d[7]:.0f} epi:acc_full={d[8]:.0f} "
              f"ph:ldtm={d[9]:.0f} ph:stg={d[10]:.0f} ph:math={d[11]:.0f} ph:fence={d[12]:.0f} ph:store={d[13]:.0f}")
        print(f"     stamps: prologue={d[20]:.0f} pdl_done={d[21]:.0f} first_acc={d[22]:.0f} epi_done={d[23]:.0f} "
              f"drained={d[24]:.0f} a_issued={d[26]:.0f} a_landed={d[27]:.0f} w_landed={d[28]:.0f}")


if __name__ == "__main__":
    B, L_in, W, Fw, Fp = 64, 3090, 25, 64, 64
    x = F.one_hot(torch.randint(0, 4, (B, L_in)), 4).to(torch.uint8).to(dev)
    w = (torch.randn((W, 4, Fw)) * 0.3).to(dev)
    bias = torch.zeros((Fw,), device=dev)
    L_out = L_in - W + 1
    out = ops.new_planes(B, L_out, Fp, False, dev)
    mask = torch.zeros((B, L_out, 1), dtype=torch.int64, device=dev)
    x8 = ops.new_x8(B, L_in, dev)
    ops.onehot_expand(x, x8)
    planes = 2
    w_op = ops.input_conv_wop(W, Fp, planes, dev)
    ops.prep_input_conv_weights(w, Fp, w_op, planes)
    ops.input_conv_fwd(x8, w_op, planes, W, Fw, bias, Fp, out, mask_out=mask)
    torch.cuda.synchronize()
    b0 = snapshot()
    N = 10
    for _ in range(N):
        ops.input_conv_fwd(x8, w_op, planes, W, Fw, bias, Fp, out, mask_out=mask)
        torch.cuda.synchronize()
    report("input_conv_fwd C1 (isolated launches)", b0, snapshot(), N)
