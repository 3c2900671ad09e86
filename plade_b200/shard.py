"""Hypothesis sharding for multi-GPU verification.

The iterations of the verification loop (PLADE/plade.cpp:547-564) are independent, so rank r verifies
hypotheses h with h % world == r on its replica of the down-sampled clouds and the ranks agree on the
winner with ONE all-reduce(MAX) of a packed 64-bit key {value bits, ~index}: the highest value wins and
ties go to the lowest hypothesis index.  `value` is an inlier count or the bit pattern of a
non-negative float score (whose bit order equals its numeric order).
"""
import numpy as np

_MASK = 0xFFFFFFFF


def shard_indices(n_hyp, rank, world):
    return np.arange(rank, n_hyp, world)


def float_bits(x):
    return int(np.float32(x).view(np.uint32))


def pack_key(value_u32, index):
    return (int(value_u32) << 32) | (_MASK - int(index))


def unpack_key(key):
    return int(key) >> 32, _MASK - (int(key) & _MASK)


def local_best_key(values_u32, indices):
    """Best (value, lowest index) of this rank's shard; 0 when the shard is empty."""
    if len(indices) == 0:
        return 0
    values_u32 = np.asarray(values_u32).astype(np.int64)
    indices = np.asarray(indices).astype(np.int64)
    best = values_u32.max()
    return pack_key(best, indices[values_u32 == best].min())


def allreduce_max_key(key, world, device=None):
    """torch.distributed all-reduce(MAX) of one int64 (NCCL on GPU tensors, gloo on CPU tensors)."""
    if world <= 1:
        return int(key)
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(key)], dtype=torch.int64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item())
