"""Read-wise sharding of a job over GPUs/ranks.

Reads are independent units: rank r of W takes a contiguous range of GLOBAL read indices.  Because every Philox
draw is addressed by the global read index, the union of the shards is bit-identical to a single-GPU run; the only
cross-read quantity, the SLOW5 aux field start_time (= samples emitted before the read, reference src/sim.c:602-604),
is an exclusive prefix sum over read lengths that the host assembles from per-shard totals.
"""
import numpy as np


def shard_range(n_reads, rank, world):
    """[lo, hi) of global read indices owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n_reads, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def start_times(len_raw_signal_per_shard):
    """Exclusive prefix sums of read lengths across shards given in rank order -> per-shard start_time arrays."""
    out, acc = [], 0
    for lens in len_raw_signal_per_shard:
        lens = np.asarray(lens, dtype=np.int64)
        out.append(acc + np.concatenate(([0], np.cumsum(lens)[:-1])) if len(lens) else lens)
        acc += int(lens.sum())
    return out
