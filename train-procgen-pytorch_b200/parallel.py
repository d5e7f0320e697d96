"""Env-sharded data parallelism: one process per GPU, envs partitioned contiguously, ONE gradient all-reduce per
optimizer step and three doubles per rollout for the global advantage moments.

The reference has no distributed code at all; this module is the whole communication layer.  It is
backend-agnostic host logic (NCCL on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend="nccl", device=None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, local_rank, world)."""
    rank, local = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if world > 1 and not dist.is_initialized():
        kw = {"device_id": torch.device(device)} if (device and backend == "nccl") else {}
        dist.init_process_group(backend, **kw)
    return rank, local, world


def shard_envs(n_envs_total, rank, world):
    """Contiguous env-index range [lo, hi) of a rank (n_envs_total must divide evenly: every rank runs the same
    kernel shapes, and per-rank minibatch = global minibatch / world keeps the optimizer-step count)."""
    if n_envs_total % world:
        raise ValueError(f"n_envs={n_envs_total} is not divisible by world_size={world}")
    per = n_envs_total // world
    return rank * per, (rank + 1) * per


def rank_seed(seed, rank, stride=1):
    """Distinct, reproducible env seed per rank (Box-World: stride = n_levels gives disjoint level ranges)."""
    return int(seed) + int(rank) * int(stride)


def shard_hyperparameters(hp, world):
    """Per-rank view of a global YAML set: n_envs and mini_batch_size are global quantities in the reference."""
    out = dict(hp)
    out["n_envs"] = hp["n_envs"] // world
    if "mini_batch_size" in hp:
        out["mini_batch_size"] = max(1, hp["mini_batch_size"] // world)
    return out


def allreduce_moments_(moments, group=None):
    """moments = [sum, sum of squares, count] (float64): after the call every rank holds the global values, so
    (adv - mean) / (std + 1e-8) equals the reference's global normalisation (common/storage.py:78-79)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(moments, op=dist.ReduceOp.SUM, group=group)
    return moments


def allreduce_gradients_(flat_grad, group=None):
    """Sum of the flat gradient over ranks (the 1/world scale is folded into the clip+Adam kernel)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad


def mean_std_from_moments(moments):
    s, ss, n = (float(x) for x in moments)
    mean = s / n
    var = max((ss - s * mean) / (n - 1.0), 0.0)
    return mean, var ** 0.5


class PeerReducer:
    """One-shot gradient all-reduce over NVLink peer memory fused with the gradient-norm reduction
    (``tpp_peer_allreduce_sqnorm``, csrc/peer_reduce.cu): replaces ``ncclAllReduce`` + ``tpp_grad_sqnorm`` in front of the
    clip + Adam kernel.  The staging buffers and signal pads are torch symmetric memory (peer-mapped allocations of
    the ranks of ONE node); the pointers go through the C-ABI as plain addresses."""

    PAD_OFFSET = 768          # our flags live behind the words torch's own symmetric-memory primitives use

    def __init__(self, n, device, group=None):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        group = group or dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise ValueError("peer all-reduce covers the GPUs of one node (<= 8)")
        self.n, self.n_pad = int(n), (int(n) + 3) // 4 * 4
        self.staging = symm_mem.empty(2 * self.n_pad, dtype=torch.float32, device=device)
        self.staging.zero_()
        self.handle = symm_mem.rendezvous(self.staging, group)
        if self.handle.signal_pad_size < 4 * (self.PAD_OFFSET + self.world):
            raise RuntimeError("symmetric-memory signal pad too small")
        self._staging_ptrs = (C.c_uint64 * self.world)(*[int(p) for p in self.handle.buffer_ptrs])
        self._pad_ptrs = (C.c_uint64 * self.world)(*[int(p) for p in self.handle.signal_pad_ptrs])
        self.reduced = torch.zeros(self.n, dtype=torch.float32, device=device)
        self.words = torch.zeros(8, dtype=torch.int32, device=device)      # [0] epoch, [2:4] tickets, [4] error flag
        torch.cuda.synchronize()
        dist.barrier(group)

    def launch(self, g_local, adam_state):
        from . import _lib
        w = self.words.data_ptr()
        _lib.call("tpp_peer_allreduce_sqnorm", self._staging_ptrs, self._pad_ptrs, self.rank, self.world,
                  self.PAD_OFFSET, _lib.ptr(g_local), _lib.ptr(self.reduced), _lib.ptr(adam_state), self.n, self.n_pad,
                  _lib.C.c_void_p(w), _lib.C.c_void_p(w + 8), _lib.C.c_void_p(w + 16), _lib.stream_ptr())

    def check(self):
        if int(self.words[4].item()) != 0:
            raise RuntimeError("peer all-reduce: a rank never raised its flag (tpp_peer_allreduce_sqnorm timed out)")
