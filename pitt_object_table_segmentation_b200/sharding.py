"""Host-side partitioning of the hot path across the GPUs of one node.

Frames and clusters are independent units: contiguous blocks per rank, no collective. One giant cloud
(config 5) is split by hypothesis: every rank holds the cloud and scores its slice of the sample
stream; the per-hypothesis counts are all-gathered (NCCL over NVLink on the GPUs, gloo in the CPU
tests) and every rank takes the earliest arg-max, which is what RandomSampleConsensus keeps
(strict '>' in ransac.hpp)."""
import numpy as np


def block_range(rank, world, n_units):
    """contiguous block [lo, hi) of n_units for `rank`; sizes differ by at most one"""
    base, rem = divmod(int(n_units), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def hypothesis_slice(rank, world, per_rank):
    """weak scaling: rank r scores stream positions [r*per_rank, (r+1)*per_rank)"""
    return rank * per_rank, (rank + 1) * per_rank


def earliest_argmax(counts):
    """index of the first maximum and its value (ties keep the earliest hypothesis)"""
    counts = np.asarray(counts)
    i = int(np.argmax(counts))  # numpy returns the first occurrence
    return i, int(counts[i])


def greedy_balance(sizes, world):
    """clusters -> ranks, largest first onto the least loaded rank (size-balanced, deterministic)"""
    order = sorted(range(len(sizes)), key=lambda i: (-sizes[i], i))
    load = [0] * world
    owner = [0] * len(sizes)
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        owner[i] = r
        load[r] += sizes[i]
    return owner
