"""2+ GPUs (torchrun): the fused peer-memory all-reduce + Adam against NCCL all-reduce + Adam from
the same state: parameters must agree bit for bit across ranks and closely between the two paths."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bpreveal_b200 import parallel  # noqa: E402
from bpreveal_b200.engine import NetConfig, SoloNet  # noqa: E402
from oracle import bpnet_oracle as O  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = dict(inputLength=52 + 300, outputLength=320, numFilters=16, numLayers=3, inputFilterWidth=3,
           outputFilterWidth=3, headTasks=[2])
cfg["inputLength"] = cfg["outputLength"] + O.length_difference(3, 3, 3)
B = 8
x = torch.tensor(O.synthetic_sequences(B * world, cfg["inputLength"], seed=5), device=dev)
c = torch.tensor(O.synthetic_profiles(B * world, cfg["outputLength"], 2, seed=5), device=dev)
xs, cs = x[rank * B:(rank + 1) * B], c[rank * B:(rank + 1) * B]


def run(use_peer, steps=6):
    net = SoloNet(NetConfig(**cfg), dev)
    net.init_random(seed=1234)
    peers = None
    if use_peer:
        peers = parallel.PeerGradExchange(net.params.flat.numel(), dev)
        net.enable_training(grads_flat=peers.grads)
    losses = []
    for s in range(steps):
        if use_peer:
            l = net.train_step(xs, cs, 0.004, peers=peers)
        else:
            l = net.train_step(xs, cs, 0.004, allreduce=lambda f: dist.all_reduce(f), world=world)
        losses.append(l.clone())
    torch.cuda.synchronize()
    p = net.params.flat.clone()
    if peers is not None:
        dist.barrier()
        peers.close()
    return p, torch.stack(losses)


p_nccl, l_nccl = run(False)
p_peer, l_peer = run(True)
gathered = [torch.empty_like(p_peer) for _ in range(world)]
dist.all_gather(gathered, p_peer)
same = all(torch.equal(gathered[0], g) for g in gathered)
d = (p_peer - p_nccl).abs().max().item()
ref = p_nccl.abs().max().item()
if rank == 0:
    print(f"world {world}: ranks bit-identical after peer exchange: {same}; "
          f"max |peer - nccl| = {d:.3e} (max |p| = {ref:.3e}); losses peer {l_peer[-1].tolist()} "
          f"nccl {l_nccl[-1].tolist()}")
    print("PEER OK" if same and d <= 1e-5 * max(ref, 1.0) else "PEER MISMATCH")
dist.destroy_process_group()
