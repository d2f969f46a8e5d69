"""Host-side plumbing for one-process-per-GPU runs.

Training is pure data parallelism: every rank runs forward/backward on an equal shard of the
global batch, the single flat fp32 gradient bucket is summed with one all-reduce (NCCL over
NVLink on GPUs; gloo in the CPU tests), scaled by 1/world inside the Adam kernel, and every rank
applies the identical update.  Prediction and attribution shard windows into contiguous ranges
with no communication; outputs are concatenated in rank order, which preserves the input order
the reference guarantees (utils.py:1221-1223, interpretUtils.py:96-99 of the reference).
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment (1 process if unset)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_distributed(backend=None, device=None):
    """Join the process group described by the environment; no-op for a single process."""
    rank, local, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl" and device is not None:
            kw["device_id"] = device
        dist.init_process_group(backend, **kw)
    return rank, local, world


def shard_range(n, rank, world):
    """Contiguous [start, stop) of rank's share of n items (window sharding for predict / PISA):
    the first n % world ranks get one extra item, concatenation in rank order restores the input."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(batch_indexes, rank, world):
    """Equal shards of one global batch.  The multinomial NLL divides by the LOCAL batch size
    (losses.py:36-38), so the average of the ranks' means is the global mean only for equal
    shards: a ragged tail (global batch not divisible by world) is dropped from every rank."""
    per = len(batch_indexes) // world
    return batch_indexes[rank * per:(rank + 1) * per]


class EpochPlan:
    """The per-epoch shuffle + jitter of H5BatchGenerator.refreshData (generators.py:108-122),
    reproduced identically on every rank: one ``default_rng(1234)`` stream, per epoch
    ``rng.shuffle(regionIndexes)`` (in place, cumulative across epochs) followed by
    ``rng.integers(0, 2 * maxJitter, numRegions, int32)``."""

    def __init__(self, numRegions, maxJitter, seed=1234):
        self.numRegions = numRegions
        self.maxJitter = maxJitter
        self.rng = np.random.default_rng(seed=seed)
        self.regionIndexes = np.arange(0, numRegions)

    def next_epoch(self):
        self.rng.shuffle(self.regionIndexes)
        sliceCols = self.rng.integers(0, self.maxJitter * 2, size=(self.numRegions,),
                                      dtype=np.int32)
        return self.regionIndexes.copy(), sliceCols


class GradAllReducer:
    """Sum of the flat gradient bucket over ranks.  ``__call__(flat)`` is what
    ``SoloNet.train_step(allreduce=...)`` expects; the 1/world scaling happens in the Adam
    kernel (``grad_scale``), not here."""

    def __init__(self, world):
        self.world = world

    def __call__(self, flat):
        if self.world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        return flat


def allreduce_scalars(values, device=None):
    """Mean over ranks of a short list of python floats (epoch loss metrics for the callbacks)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t /= dist.get_world_size()
    return [float(v) for v in t.cpu()]


def broadcast_object(obj, src=0):
    """Rank ``src`` decides (learning-rate drops, early stopping, adaptive lambda); all follow."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        box = [obj]
        dist.broadcast_object_list(box, src=src)
        return box[0]
    return obj


class PeerGradExchange:
    """Gradient buckets and flag blocks of all ranks of one node, mapped into every process over
    CUDA IPC, for the optimiser kernel that does the all-reduce itself
    (``bpb_adam_prep_step_peer``).  ``grads`` is this rank's bucket as a torch tensor: the engine
    writes its gradients there (``SoloNet.enable_training(grads_flat=exchange.grads)``).

    Needs an initialised ``torch.distributed`` group (the 64-byte IPC handles travel through
    ``all_gather_object``), all ranks on one node with peer access."""

    FLAG_BYTES = 128

    def __init__(self, numel, device, rank=None, world=None):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _lib
        lib = _lib.load()
        self.rank = rank if rank is not None else dist.get_rank()
        self.world = world if world is not None else dist.get_world_size()
        if not 1 <= self.world <= 8:
            raise ValueError("PeerGradExchange supports 1..8 ranks on one node")
        self.numel = int(numel)
        nbytes = self.FLAG_BYTES + 4 * self.numel
        base = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        _lib.check(lib.bpb_peer_alloc(nbytes, C.byref(base), handle), "bpb_peer_alloc")
        self._base = base.value
        handles = [bytes(handle)]
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle))
        self._opened = []
        bases = []
        for r in range(self.world):
            if r == self.rank:
                bases.append(self._base)
                continue
            ptr = C.c_void_p()
            hb = (C.c_ubyte * 64).from_buffer_copy(handles[r])
            _lib.check(lib.bpb_peer_open(hb, C.byref(ptr)), "bpb_peer_open")
            self._opened.append(ptr.value)
            bases.append(ptr.value)
        self.args = _lib.PeerArgs()
        self.args.world, self.args.rank = self.world, self.rank
        for r in range(self.world):
            self.args.flags[r] = bases[r]
            self.args.grads[r] = bases[r] + self.FLAG_BYTES

        class _Raw:  # zero-copy view of the bucket for torch (CUDA array interface)
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": (self.numel,), "typestr": "<f4",
                                        "data": (self._base + self.FLAG_BYTES, False),
                                        "version": 2, "strides": None}
        self._raw = raw
        self.grads = torch.as_tensor(raw, device=device)
        fraw = _Raw()
        fraw.__cuda_array_interface__ = {"shape": (self.FLAG_BYTES // 4,), "typestr": "<i4",
                                         "data": (self._base, False), "version": 2,
                                         "strides": None}
        self._fraw = fraw
        self.flags = torch.as_tensor(fraw, device=device)  # words 17: wait bound (s), 18: error
        self.set_timeout(float(os.environ.get("BPB_PEER_TIMEOUT_S", "30")))
        if self.world > 1:
            dist.barrier()  # every rank has mapped every bucket before anyone uses them

    def set_timeout(self, seconds):
        """Wall-clock bound of the in-kernel waits for a peer (default 30 s, ``BPB_PEER_TIMEOUT_S``)."""
        self.flags[17] = max(1, int(round(seconds)))

    def check(self):
        """Raises if a kernel of this rank gave up waiting for a peer (the kernels record it in
        the flag block instead of trapping, so the CUDA context stays usable and the caller can
        fall back to ``GradAllReducer``).  Synchronises: call it at epoch boundaries."""
        err = int(self.flags[18].item())
        if err:
            self.flags[18] = 0
            raise RuntimeError(
                f"rank {self.rank}: peer rank {err - 1} did not reach the gradient exchange within "
                f"{int(self.flags[17].item())} s (parameters of the ranks may have diverged)")

    def close(self):
        from . import _lib
        lib = _lib.load()
        for ptr in self._opened:
            lib.bpb_peer_close(ptr)
        self._opened = []
        if self._base:
            lib.bpb_peer_free(self._base)
            self._base = None
