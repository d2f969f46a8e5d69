"""Role profile of the tower's data gradient (library built with `make PROFILE=1`): cycles per
launch that each warp role of conv_tc_kernel spends in its mbarrier waits and epilogue phases,
for the materialised path (operand gz, outputs g and gz) and the masked-operand path (operand
g (.) bits masked in shared memory, output g).  Usage: diag_roles_bwd.py [dilation ...]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bpreveal_b200 import _lib, ops  # noqa: E402
from tests.diag_kernels import prep_w, split  # noqa: E402

dev = torch.device("cuda:0")
NAMES = {0: "masker:fence+arrive", 7: "masker:work", 1: "producer:empty", 2: "producer:rempty", 3: "mma:w_full", 4: "mma:acc_empty",
         5: "mma:a_ready", 6: "epi:rfull", 8: "epi:acc_full", 9: "ph:acc+ldtm",
         10: "ph:stg_free", 11: "ph:math", 12: "ph:fence", 13: "ph:bar+store", 14: "masker:full"}


def snap():
    buf = (C.c_uint * 2048)()
    _lib.load().bpb_debug_dump(buf, 2048)
    return list(buf)


def run(label, fn, N=10):
    fn()
    torch.cuda.synchronize()
    b0 = snap()
    for _ in range(N):
        fn()
    torch.cuda.synchronize()
    b1 = snap()
    for cta in range(2):
        base = 1024 + 32 * cta
        tot = (b1[base + 15] - b0[base + 15]) / N
        parts = [f"{nm}={(b1[base + g] - b0[base + g]) / N:.0f}" for g, nm in NAMES.items()
                 if b1[base + g] != b0[base + g]]
        print(f"  {label} cta{cta}: total={tot:.0f}cyc  " + "  ".join(parts))


B, Fn, L = 64, 64, 3066
for d in [int(a) for a in sys.argv[1:]] or [2, 64, 256]:
    Lg = L - 2 * d
    g = split(torch.randn((B, Lg, Fn), device=dev), False)
    w = torch.randn((3, Fn, Fn), device=dev) / 14
    _, bwd = prep_w(w, Fn, False)
    out = ops.new_planes(B, L, Fn, False, dev)
    out2 = ops.new_planes(B, L, Fn, False, dev)
    mask_in = torch.randint(-2 ** 62, 2 ** 62, (B, L, 1), dtype=torch.int64, device=dev)
    mask_a = torch.randint(-2 ** 62, 2 ** 62, (B, Lg, 1), dtype=torch.int64, device=dev)
    print(f"conv dgrad d={d} (B={B}, L={L}): cycles per launch, summed over the lanes 0 of a role")
    if not os.environ.get("ROLES_ONLY_XF"):
      run("materialised", lambda: ops.conv3(g, bwd, (0, -d, -2 * d), 1, L, resid=g, resid_off=-d,
                                            out=out, out2=out2, mask_in=mask_in))
    run("masked-operand", lambda: ops.conv3(g, bwd, (0, -d, -2 * d), 1, L, resid=g, resid_off=-d,
                                            out=out, mask_a=mask_a))
