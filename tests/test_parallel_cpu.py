"""CPU (gloo, world_size 2): the data-parallel host logic of bpreveal_b200.parallel.

The gradient "kernels" here are the fp64 oracle on tiny nets -- what is under test is the sharding
arithmetic (equal shards, one flat bucket, sum then 1/world, identical update on every rank) and
the order-preserving window sharding of predict / PISA."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bpreveal_b200 import parallel
from oracle import bpnet_oracle as O
from tests import helpers as Hh


def test_shard_range_preserves_order_and_covers():
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_batch_equal_and_disjoint():
    idx = np.arange(64)
    parts = [parallel.shard_batch(idx, r, 4) for r in range(4)]
    assert all(len(p) == 16 for p in parts)
    assert np.array_equal(np.concatenate(parts), idx)
    # ragged global batch: every rank drops the tail so that shards stay equal
    parts = [parallel.shard_batch(np.arange(10), r, 4) for r in range(4)]
    assert [len(p) for p in parts] == [2, 2, 2, 2]


def test_epoch_plan_matches_reference_generator_semantics():
    """generators.py:108-122: shuffle is in place and cumulative; both calls come from one
    default_rng(1234) stream in this order."""
    plan = parallel.EpochPlan(50, 5)
    rng = np.random.default_rng(seed=1234)
    reg = np.arange(50)
    for _ in range(3):
        got_r, got_c = plan.next_epoch()
        rng.shuffle(reg)
        want_c = rng.integers(0, 10, size=(50,), dtype=np.int32)
        assert np.array_equal(got_r, reg) and np.array_equal(got_c, want_c)
        want_r, want_j = O.epoch_jitter(50, 5, np.random.default_rng(seed=1234))
    # every rank builds the identical plan
    a, b = parallel.EpochPlan(31, 3), parallel.EpochPlan(31, 3)
    for _ in range(2):
        ra, ca = a.next_epoch()
        rb, cb = b.next_epoch()
        assert np.array_equal(ra, rb) and np.array_equal(ca, cb)


def _flat_grads(params, x, counts):
    tot, pl, cl, grads = Hh.oracle_loss_and_grads(params, x, counts, [1.0], [1.0])
    keys = sorted(grads)
    return torch.tensor(np.concatenate([grads[k].ravel() for k in keys])), tot


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, _, w = parallel.init_distributed(backend="gloo")
    assert (r, w) == (rank, world)
    cfg = Hh.tiny_cfg()
    params = Hh.oracle_params(cfg)
    B = 8
    x = O.synthetic_sequences(B, cfg["inputLength"], seed=3)
    counts = O.synthetic_profiles(B, cfg["outputLength"], 2, seed=3, meanReads=30.0)
    mine = parallel.shard_batch(np.arange(B), rank, world)
    flat, tot = _flat_grads(params, x[mine], counts[mine])
    parallel.GradAllReducer(world)(flat)
    flat /= world
    mean_loss = parallel.allreduce_scalars([tot])[0]
    decision = parallel.broadcast_object({"lr": 0.002, "stop": False} if rank == 0 else None)
    lo, hi = parallel.shard_range(11, rank, world)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), flat=flat.numpy(), loss=mean_loss,
             lr=decision["lr"], lo=lo, hi=hi)
    dist.destroy_process_group()


def test_gloo_two_ranks_match_single_process(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    outs = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    cfg = Hh.tiny_cfg()
    params = Hh.oracle_params(cfg)
    x = O.synthetic_sequences(8, cfg["inputLength"], seed=3)
    counts = O.synthetic_profiles(8, cfg["outputLength"], 2, seed=3, meanReads=30.0)
    want, tot = _flat_grads(params, x, counts)
    for o in outs:
        np.testing.assert_allclose(o["flat"], want.numpy(), rtol=1e-10, atol=1e-12)
        assert abs(float(o["loss"]) - tot) < 1e-9 * abs(tot)
        assert float(o["lr"]) == 0.002
    assert np.array_equal(outs[0]["flat"], outs[1]["flat"])  # identical update on every rank
    assert (int(outs[0]["lo"]), int(outs[0]["hi"]), int(outs[1]["lo"]), int(outs[1]["hi"])) == \
        (0, 6, 6, 11)
