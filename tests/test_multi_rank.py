"""N>1 host logic on CPU: two gloo ranks shard a job read-wise, each generates its shard (through the oracle here,
through libsqg.so on GPUs), and the union equals the single-rank run; start_time is assembled from shard totals."""
import hashlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import helpers as H
from squigulator_b200.shard import shard_range, start_times, CpgIndex


def test_shard_ranges_partition():
    for n in (0, 1, 7, 8, 1000, 1001):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def _worker(rank, world, port, n_reads, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prof, flags = H.PRESETS["dna-r10-prom"]
    # rank 0 owns the pore-model table and broadcasts it once (NCCL on GPUs, gloo here); no other collective on the path
    model = torch.zeros(2 * 4 ** 9, dtype=torch.float32)
    if rank == 0:
        model.copy_(torch.from_numpy(H.random_model(4 ** 9)))
    dist.broadcast(model, src=0)
    reads = H.random_reads(n_reads, 800, seed=4)      # every rank can see the read list; it generates only its range
    lo, hi = shard_range(n_reads, rank, world)
    o = H.Oracle(H.load_oracle(), prof, flags, 9, 4 ** 9, model.numpy(), 5, H.RNG_PHILOX, ztable=H.load_ztable())
    digests, lens = [], []
    for g in range(lo, hi):
        r = o.gen_sig(reads[g], read_index=g)
        digests.append(hashlib.sha256(r["sig"].tobytes()).digest())
        lens.append(len(r["sig"]))
    o.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, (digests, lens))
    if rank == 0:
        ret.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one_rank():
    n_reads = 11
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_reads, ret)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    digests = [d for shard in gathered for d in shard[0]]
    lens = [shard[1] for shard in gathered]
    # single-rank run
    prof, flags = H.PRESETS["dna-r10-prom"]
    o = H.Oracle(H.load_oracle(), prof, flags, 9, 4 ** 9, H.random_model(4 ** 9), 5, H.RNG_PHILOX, ztable=H.load_ztable())
    reads = H.random_reads(n_reads, 800, seed=4)
    whole = [o.gen_sig(reads[g], read_index=g) for g in range(n_reads)]
    o.close()
    assert digests == [hashlib.sha256(w["sig"].tobytes()).digest() for w in whole]
    st = np.concatenate(start_times(lens))
    exp = np.concatenate(([0], np.cumsum([len(w["sig"]) for w in whole])[:-1]))
    assert np.array_equal(st, exp)


# ---- methylated coordinate batches shard too: every rank's position in the reference's rand_meth stream comes from a
# host-side CpG prefix index (squigulator_b200.shard.CpgIndex), so no rank waits for another one's meth_draws ----

def _coord_list(contigs, n, seed):
    rs = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        c = int(rs.randint(0, len(contigs)))
        ln = int(min(len(contigs[c]), max(0, rs.gamma(2.0, 300))))
        out.append((c, int(rs.randint(0, len(contigs[c]) - ln + 1)), ln, "+-"[int(rs.randint(0, 2))]))
    return out


def test_cpg_index_counts_equal_oracle_draws():
    import ctypes as C
    lib = H.load_oracle()
    contigs, marr = H.synthetic_genome(seed=5, n_contigs=3, mean_len=2000)
    idx = CpgIndex(contigs)
    st = C.c_int64(7)
    for c, pos, ln, strand in _coord_list(contigs, 200, seed=1) + [(0, 0, 0, "+"), (0, 5, 1, "-"), (1, 0, len(contigs[1]), "-")]:
        _, d = H.oracle_extract_read(lib, contigs[c], marr[c], pos, ln, strand, st)
        assert idx.draws(c, pos, ln) == d
    # contigs without methylation data take no draws
    assert CpgIndex(contigs, contig_has_meth=[0, 1, 0]).draws(0, 0, len(contigs[0])) == 0


def _meth_worker(rank, world, port, n_reads, seed, ret):
    import ctypes as C
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = H.load_oracle()
    contigs, marr = H.synthetic_genome(seed=9, n_contigs=4, mean_len=2500)
    coords = _coord_list(contigs, n_reads, seed=3)
    lo, hi = shard_range(n_reads, rank, world)
    # this rank's position in rand_meth: draws of all reads before its shard (no communication needed; the all-gather
    # below only cross-checks the per-shard totals the way a driver would)
    idx = CpgIndex(contigs)
    bases = idx.draw_bases(coords)
    st = C.c_int64(lib.sqo_lehmer_jump(seed + 6, int(bases[lo])))
    digests, draws = [], 0
    for g in range(lo, hi):
        c, pos, ln, strand = coords[g]
        b, d = H.oracle_extract_read(lib, contigs[c], marr[c], pos, ln, strand, st)
        digests.append(hashlib.sha256(b).digest())
        draws += d
    totals = [None] * world
    dist.all_gather_object(totals, draws)
    assert sum(totals[:rank]) == int(bases[lo]) and sum(totals) == int(bases[-1])
    gathered = [None] * world
    dist.all_gather_object(gathered, digests)
    if rank == 0:
        ret.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_methylated_coordinates_equal_one_rank():
    import ctypes as C
    n_reads, seed = 37, 12
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_meth_worker, args=(r, 2, port, n_reads, seed, ret)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lib = H.load_oracle()
    contigs, marr = H.synthetic_genome(seed=9, n_contigs=4, mean_len=2500)
    st = C.c_int64(seed + 6)   # one rank: the stream simply runs through all reads in order, as in `squigulator -t1`
    whole = []
    for c, pos, ln, strand in _coord_list(contigs, n_reads, seed=3):
        whole.append(hashlib.sha256(H.oracle_extract_read(lib, contigs[c], marr[c], pos, ln, strand, st)[0]).digest())
    assert [d for shard in gathered for d in shard] == whole
