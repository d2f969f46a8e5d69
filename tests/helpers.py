"""Shared helpers for the parity tests (oracle side)."""
import numpy as np
import torch

from oracle import bpnet_oracle as O


def tiny_cfg(**kw):
    """F=8, 3 layers, 3/3, 20 out -> 52 in: the reference's own lengthCalc example
    (doc/demos/osknExample.ipynb cell 88)."""
    d = dict(inputLength=52, outputLength=20, numFilters=8, numLayers=3, inputFilterWidth=3,
             outputFilterWidth=3, headTasks=[2])
    d.update(kw)
    return d


def mid_cfg(**kw):
    nl, wi, wo, lo = 5, 25, 23, 300
    d = dict(inputLength=lo + O.length_difference(nl, wi, wo), outputLength=lo, numFilters=64,
             numLayers=nl, inputFilterWidth=wi, outputFilterWidth=wo, headTasks=[2])
    d.update(kw)
    if "inputLength" not in kw:
        d["inputLength"] = d["outputLength"] + O.length_difference(
            d["numLayers"], d["inputFilterWidth"], d["outputFilterWidth"])
    return d


def oracle_params(cfg, seed=1234, scale=1.0, biasScale=0.1):
    return O.init_solo_params(cfg["numFilters"], cfg["numLayers"], cfg["inputFilterWidth"],
                              cfg["outputFilterWidth"], cfg["headTasks"], seed=seed, scale=scale,
                              biasScale=biasScale)


def oracle_forward(params, x_u8, dtype=torch.float64):
    p = O.to_torch(params, dtype)
    x = torch.tensor(np.asarray(x_u8), dtype=dtype)
    profs, cnts = O.solo_forward(x, p)
    return torch.cat(profs, dim=2).numpy(), torch.cat(cnts, dim=1).numpy()


def oracle_loss_and_grads(params, x_u8, counts, profileWeights, lambdas, dtype=torch.float64):
    """Returns (total, mnll list, wmse list, grads dict) following training.py:64-65."""
    p = O.to_torch(params, dtype, requires_grad=True)
    x = torch.tensor(np.asarray(x_u8), dtype=dtype)
    profs, cnts = O.solo_forward(x, p)
    k = 0
    trueP, trueC = [], []
    ct = torch.tensor(np.asarray(counts), dtype=dtype)
    for h, pr in enumerate(profs):
        T = pr.shape[2]
        trueP.append(ct[:, :, k:k + T])
        trueC.append(torch.log(ct[:, :, k:k + T].sum(dim=(1, 2))))
        k += T
    tot, pl, cl = O.total_loss(profs, cnts, trueP, trueC, profileWeights, lambdas)
    tot.backward()
    grads = {k_: v.grad.numpy() for k_, v in p.items()}
    return tot.item(), [v.item() for v in pl], [v.item() for v in cl], grads


def rel_err(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.abs(got - ref).max() / (np.abs(ref).max() + 1e-30))
