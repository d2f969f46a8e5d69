"""Role/timeline profile (library built with `make PROFILE=1`) of the split-K weight-gradient
kernel for the three ways the C1 step uses it, plus device times of the small edge kernels."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bpreveal_b200 import _lib, ops  # noqa: E402
from bpreveal_b200.engine import NetConfig, SoloNet  # noqa: E402
from oracle import bpnet_oracle as O  # noqa: E402

C1 = dict(inputLength=3090, outputLength=1000, numFilters=64, numLayers=9, inputFilterWidth=25,
          outputFilterWidth=23, headTasks=[2])
dev = torch.device("cuda:0")


def snapshot():
    buf = (C.c_uint * 2048)()
    _lib.load().bpb_debug_dump(buf, 2048)
    return buf


def profile(label, fn, n=10):
    fn()
    torch.cuda.synchronize()
    b0 = snapshot()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    tot = 0.0
    for _ in range(n):
        ev[0].record()
        fn()
        ev[1].record()
        torch.cuda.synchronize()
        tot += ev[0].elapsed_time(ev[1])
    b1 = snapshot()
    print(f"{label}: {tot / n * 1000:.1f} us per call (isolated, events)")
    for cta in range(2):
        base = 1024 + 32 * cta
        d = [(b1[base + i] - b0[base + i]) / n for i in range(32)]
        print(f"   cta{cta}: total={d[15]:.0f} visits={d[16]:.0f} prologue={d[20]:.0f} pdl_done={d[21]:.0f} "
              f"first_full={d[22]:.0f} mma_done={d[23]:.0f} acc_ready={d[24]:.0f}")


net = SoloNet(NetConfig(**C1), dev)
net.init_random(seed=1234)
B = 64
x = torch.tensor(O.synthetic_sequences(B, 3090, seed=1), device=dev)
c = torch.tensor(O.synthetic_profiles(B, 1000, 2, seed=1), device=dev)
net.train_step(x, c, 0.004, use_graph=False)
torch.cuda.synchronize()
ws, cfg = net._workspace(B, "train"), net.cfg
g = net.g
Ls = cfg.lengths()
nL = cfg.numLayers
planes = 1
aL = ws["acts"][nL]
profile("input_conv_wgrad", lambda: ops.input_conv_wgrad(
    ws["x8"], ws["gzl"][0], cfg.inputFilterWidth, cfg.numFilters, g["conv1_w"], g["conv1_b"],
    ws["scratch"]))
profile("heads_wgrad", lambda: ops.heads_wgrad(
    aL, ws["d8"], cfg.outputFilterWidth, cfg.numFilters, cfg.T, cfg.H, net.cp, planes,
    g["prof_w"], g["cnt_w"], ws["scratch"]))
profile("wgrad3 layer 0 (d=2)", lambda: ops.wgrad3(
    ws["acts"][0], ws["gzl"][1], 2, g["dil_w"][0], g["dil_b"][0], ws["scratch"]))
profile("wgrad3_multi (9 layers)", lambda: ops.wgrad3_multi(
    [(ws["acts"][i], ws["gzl"][i + 1], 2 ** (i + 1), g["dil_w"][i], g["dil_b"][i])
     for i in range(nL)], ws["scratch"]))


def timed(label, fn, n=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(n):
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    print(f"{label}: {tot / n * 1000:.1f} us per call (isolated, events)")


timed("heads_fwd_tc", lambda: ops.heads_fwd_tc(
    aL, net.w_hf_op, cfg.outputFilterWidth, cfg.T, cfg.H, net.p["prof_b"], net.p["cnt_b"],
    ws["logits"], ws["logcounts"], ws["hf_ws"]))
timed("loss_fwd_bwd", lambda: ops.loss_fwd_bwd(
    net.task_off, ws["logits"], ws["logcounts"], ws["counts"], net.profile_weight, net.lam,
    ws["losses"], ws["dlogits"], ws["dlogcounts"]))
timed("heads_pack_grad(+bias grads)", lambda: ops.heads_pack_grad(
    ws["dlogits"], ws["dlogcounts"], Ls[nL], cfg.outputFilterWidth, net.cp, planes, ws["d8"],
    dpb=g["prof_b"], dcb=g["cnt_b"]))
timed("heads_pack_grad (no bias grads)", lambda: ops.heads_pack_grad(
    ws["dlogits"], ws["dlogcounts"], Ls[nL], cfg.outputFilterWidth, net.cp, planes, ws["d8"]))
timed("heads_dgrad", lambda: ops.heads_dgrad(
    ws["d8"], net.w_hd_op, B, Ls[nL], cfg.outputFilterWidth, cfg.Fp, net.cp, planes,
    g=net._sub(ws["g"][nL % 2], Ls[nL]), gz=ws["gzl"][nL], mask_in=ws["masks"][nL]))
timed("input_conv_fwd", lambda: ops.input_conv_fwd(
    ws["x8"], net.w_in_op, net.w_in_planes, cfg.inputFilterWidth, cfg.numFilters,
    net.p["conv1_b"], cfg.Fp, ws["acts"][0], mask_out=ws["masks"][0]))

b0 = snapshot()
for _ in range(10):
    ops.loss_fwd_bwd_pack(net.task_off, net.host_task_off, ws["logits"], ws["logcounts"], ws["counts"],
                          net.profile_weight, net.lam, ws["losses"], Ls[nL], cfg.outputFilterWidth,
                          net.cp, planes, ws["d8"], ws["loss_ws"], dlogits=ws["dlogits"],
                          dlogcounts=ws["dlogcounts"], dpb=g["prof_b"], dcb=g["cnt_b"])
    torch.cuda.synchronize()
b1 = snapshot()
print("loss_pack block 0 stamps (cycles):", [(b1[1024 + i] - b0[1024 + i]) // 10 for i in range(1, 10)])


b0 = snapshot()
for _ in range(10):
    ops.heads_fwd_tc(aL, net.w_hf_op, cfg.outputFilterWidth, cfg.T, cfg.H, net.p["prof_b"],
                     net.p["cnt_b"], ws["logits"], ws["logcounts"], ws["hf_ws"])
    torch.cuda.synchronize()
b1 = snapshot()
for cta in range(2):
    base = 1024 + 32 * cta
    d = [(b1[base + i] - b0[base + i]) // 10 for i in range(32)]
    print(f"heads_fwd_tc cta{cta}: total={d[15]} prologue={d[20]} w_issued={d[21]} w_landed={d[22]} "
          f"first_full={d[23]} first_mma_issued={d[24]} mma_all_issued={d[25]} first_acc={d[26]}")
