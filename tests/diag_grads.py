"""Debug aid: per-parameter gradient errors of the GPU path against the fp64 oracle."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import bpnet_oracle as O  # noqa: E402
from tests import helpers as Hh  # noqa: E402
from tests.test_gpu_model import CASES, build  # noqa: E402

dev = torch.device("cuda:0")
for name, cfg in CASES:
    for precise in (False, True):
        net, params = build(cfg, precise, dev, scale=1.5)
        net.enable_training()
        B = 6
        H = len(cfg["headTasks"])
        T = sum(cfg["headTasks"])
        x = O.synthetic_sequences(B, cfg["inputLength"], seed=11)
        counts = O.synthetic_profiles(B, cfg["outputLength"], T, seed=11, meanReads=50.0)
        pw = [1.0 + 0.5 * h for h in range(H)]
        lam = [2.0 + h for h in range(H)]
        net.profile_weight.copy_(torch.tensor(pw))
        net.lam.copy_(torch.tensor(lam))
        tot, pl, cl, grads = Hh.oracle_loss_and_grads(params, x, counts, pw, lam)
        ws = net._workspace(B, "train")
        ws["x"].copy_(torch.tensor(x, device=dev))
        ws["counts"].copy_(torch.tensor(counts, device=dev))
        net.forward_backward_into(ws)
        torch.cuda.synchronize()
        got = net.state(net.grads_flat)
        gscale = max(np.abs(v).max() for v in grads.values())
        errs = {k: np.abs(got[k] - grads[k]).max() / max(np.abs(grads[k]).max(), 1e-3 * gscale)
                for k in grads}
        print(name, "precise" if precise else "bf16",
              " ".join(f"{k}={v:.1e}" for k, v in errs.items()), flush=True)
