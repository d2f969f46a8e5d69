"""Host vs device time of the end-to-end training call (host inputs)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bpreveal_b200.engine import NetConfig, SoloNet  # noqa: E402
from oracle import bpnet_oracle as O  # noqa: E402

C1 = dict(inputLength=3090, outputLength=1000, numFilters=64, numLayers=9, inputFilterWidth=25,
          outputFilterWidth=23, headTasks=[2])
dev = torch.device("cuda:0")
net = SoloNet(NetConfig(**C1), dev)
net.init_random(seed=1234)
B = 64
xh = torch.tensor(O.synthetic_sequences(4 * B, 3090, seed=1)).pin_memory()
ch = torch.tensor(O.synthetic_profiles(4 * B, 1000, 2, seed=1)).pin_memory()
xd, cd = xh.to(dev), ch.to(dev)
loss_host = torch.zeros((2,), dtype=torch.float32).pin_memory()
for mode in ("device", "host", "host"):
    xs, cs = (xd, cd) if mode == "device" else (xh, ch)
    for i in range(5):
        net.train_step(xs[:B], cs[:B], 0.004)
    torch.cuda.synchronize()
    n = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(n):
        j = i % 4
        losses = net.train_step(xs[j * B:(j + 1) * B], cs[j * B:(j + 1) * B], 0.004)
        loss_host.copy_(losses, non_blocking=True)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{mode}: host enqueue {1e3 * (t1 - t0) / n:.3f} ms/step, wall {1e3 * (t2 - t0) / n:.3f}, "
          f"device {e0.elapsed_time(e1) / n:.3f} ms/step")
