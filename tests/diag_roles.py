"""Role profile (library built with `make PROFILE=1`): where each warp role of the conv kernel
waits.  Prints, for CTA 0..3, cycles spent in every mbarrier wait group."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bpreveal_b200 import _lib, ops  # noqa: E402
from tests.diag_kernels import prep_w, split  # noqa: E402
import ctypes as C  # noqa: E402

dev = torch.device("cuda:0")
NAMES = {1: "producer:empty", 2: "producer:rempty", 3: "mma:w_full", 4: "mma:acc_empty",
         5: "mma:full", 6: "epi:rfull", 7: "epi:full(halo)", 8: "epi:acc_full", 9: "ph:acc+ldtm",
         10: "ph:stg_free", 11: "ph:math", 12: "ph:fence", 13: "ph:bar+store", 15: "TOTAL"}


def dump(label, launches):
    buf = (C.c_uint * 2048)()
    _lib.load().bpb_debug_dump(buf, 2048)
    print(label)
    for cta in range(2):
        base = 1024 + 32 * cta
        tot = buf[base + 15] / launches
        parts = []
        for g, nm in NAMES.items():
            if g == 15 or buf[base + 16 + g] == 0:
                continue
            parts.append(f"{nm}={buf[base + g] / launches:.0f}cyc/{buf[base + 16 + g] // launches}w")
        print(f"  cta{cta}: total={tot:.0f}  " + "  ".join(parts))


B, Fn, L, d = 64, 64, 3066, int(sys.argv[1]) if len(sys.argv) > 1 else 2
x = torch.randn((B, L, Fn), device=dev)
w = torch.randn((3, Fn, Fn), device=dev) / 14
xin = split(x, False)
fwd, bwd = prep_w(w, Fn, False)
L_out = L - 2 * d
out = ops.new_planes(B, L_out, Fn, False, dev)
mask = torch.zeros((B, L_out, 1), dtype=torch.int64, device=dev)
bias = torch.zeros((Fn,), device=dev)
ops.conv3(xin, fwd, (0, d, 2 * d), 0, L_out, bias=bias, resid=xin, resid_off=d, out=out, mask_out=mask)
torch.cuda.synchronize()
# reset counters by reading baseline
buf0 = (C.c_uint * 2048)()
_lib.load().bpb_debug_dump(buf0, 2048)
N = 10
for _ in range(N):
    ops.conv3(xin, fwd, (0, d, 2 * d), 0, L_out, bias=bias, resid=xin, resid_off=d, out=out, mask_out=mask)
    if os.environ.get("ROLES_SYNC"):
        torch.cuda.synchronize()
torch.cuda.synchronize()
buf1 = (C.c_uint * 2048)()
_lib.load().bpb_debug_dump(buf1, 2048)
print(f"conv fwd d={d}: per launch, CTA 0 (waits are summed over the lanes-0 of all warps of a role)")
for cta in range(2):
    base = 1024 + 32 * cta
    tot = (buf1[base + 15] - buf0[base + 15]) / N
    parts = []
    for g, nm in NAMES.items():
        if g == 15:
            continue
        n = (buf1[base + 16 + g] - buf0[base + 16 + g]) // N
        if n == 0 and g < 9:
            continue
        parts.append(f"{nm}={(buf1[base + g] - buf0[base + g]) / N:.0f}cyc/{n}waits")
    print(f"  cta{cta}: total={tot:.0f}cyc  " + "  ".join(parts))
    st = [(buf1[base + s] - buf0[base + s]) / N for s in range(20, 31)]
    print(f"     stamps: prologue_done={st[0]:.0f} pdl_wait_done={st[1]:.0f} first_acc={st[2]:.0f} "
          f"epi_done={st[3]:.0f} stores_drained={st[4]:.0f} w_issued={st[5]:.0f} a_issued={st[6]:.0f} a_landed={st[7]:.0f} w_landed={st[8]:.0f} loop_entered={st[9]:.0f} expect_done={st[10]:.0f}")
