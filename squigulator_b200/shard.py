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


class CpgIndex:
    """Host-side prefix index of CpG sites per contig, for coordinate batches with methylation (sqg_submit_coords).

    The reference takes one value of its rand_meth stream per forward-strand `CG` inside a read
    (methylate_dna, reference src/genread.c:207-241: contig[pos+i] == 'C', contig[pos+i+1] == 'G', i+1 < rlen), on
    contigs that have methylation data.  The number of draws of a read is therefore a difference of prefix counts, so
    `meth_draw_base` of every batch and of every rank's shard can be computed up front, without waiting for the
    previous batch's `meth_draws`, and methylated jobs shard like all others."""

    def __init__(self, contigs, contig_has_meth=None):
        self.prefix = []
        for c, seq in enumerate(contigs):
            a = np.frombuffer(seq, dtype=np.uint8)
            site = np.zeros(len(a) + 1, dtype=np.int64)
            if len(a) > 1 and (contig_has_meth is None or contig_has_meth[c]):
                site[1:len(a)] = (a[:-1] == ord("C")) & (a[1:] == ord("G"))
            self.prefix.append(np.cumsum(site))  # prefix[p] = sites at positions < p

    def draws(self, contig, pos, length):
        """rand_meth draws of one read: sites at positions pos .. pos+length-2"""
        if length < 2:
            return 0
        p = self.prefix[contig]
        return int(p[pos + length - 1] - p[pos])

    def draw_bases(self, coords, base=0):
        """exclusive prefix of the draws over reads given as (contig, pos, len, strand): entry i = meth_draw_base of a
        batch that starts at read i; the last entry is the stream position after all reads"""
        out = np.empty(len(coords) + 1, dtype=np.int64)
        acc = base
        for i, (c, pos, ln, _s) in enumerate(coords):
            out[i] = acc
            acc += self.draws(c, pos, ln)
        out[len(coords)] = acc
        return out
