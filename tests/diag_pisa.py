"""PISA throughput against the batch of query bases per CUDA-graph replay, and a per-family device
time breakdown of one replay."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bpreveal_b200 import interpret, ops  # noqa: E402
from bpreveal_b200.engine import NetConfig, SoloNet  # noqa: E402
from oracle import bpnet_oracle as O  # noqa: E402

C1 = dict(inputLength=3090, outputLength=1000, numFilters=64, numLayers=9, inputFilterWidth=25,
          outputFilterWidth=23, headTasks=[2])
dev = torch.device("cuda:0")
net = SoloNet(NetConfig(**C1), dev)
net.init_random(seed=1234)
x = torch.tensor(O.synthetic_sequences(96, 3090, seed=1), device=dev)
for Q in [int(a) for a in sys.argv[1:]] or [6, 12, 24]:
    interpret.pisa_batch(net, x[:Q], headID=0, taskID=0, numShuffles=20)
    torch.cuda.synchronize()
    reps = max(2, 48 // Q)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        interpret.pisa_batch(net, x[(r * Q) % 72:(r * Q) % 72 + Q], headID=0, taskID=0, numShuffles=20)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"Q={Q}: {ms:.3f} ms per replay, {Q / ms * 1e3:.0f} bases/s", flush=True)

# per-family breakdown of one un-graphed evaluation (Q = 6)
calls = []
names = ["conv3", "input_conv_fwd", "input_conv_dgrad", "heads_fwd", "heads_fwd_tc", "heads_dgrad",
         "heads_pack_grad", "onehot_expand", "shap_combine", "shuffle_rows", "flat_profile_seed"]
real = {}
for nm in names:
    if not hasattr(ops, nm):
        continue
    fn = getattr(ops, nm)
    real[nm] = fn

    def mk(nm, fn):
        def inner(*a, **k):
            key = nm
            if nm == "conv3":
                key = f"conv3_mode{a[3]}" + ("_mult" if k.get("mult") is not None else "") + \
                      ("_zx" if k.get("zx_in") is not None else "") + ("_z" if k.get("z_out") is not None else "")
            calls.append((key, fn, a, k))
            return fn(*a, **k)
        return inner
    setattr(ops, nm, mk(nm, fn))
Q = 6
interpret.pisa_batch(net, x[:Q], headID=0, taskID=0, numShuffles=20, use_graph=False)
calls.clear()
interpret.pisa_batch(net, x[:Q], headID=0, taskID=0, numShuffles=20, use_graph=False)
torch.cuda.synchronize()
for nm, fn in real.items():
    setattr(ops, nm, fn)
keys = list(dict.fromkeys(c[0] for c in calls))
tot = 0.0
for key in keys:
    fam = [c for c in calls if c[0] == key]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _, fn, a, k in fam:
            fn(*a, **k)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 10
    tot += t
    print(f"  {key:28s} {len(fam):3d} launches {t * 1e3:8.1f} us")
print(f"  total {tot * 1e3:.1f} us for {Q} bases")
