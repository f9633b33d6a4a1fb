#!/usr/bin/env python3
"""bench.py — int16 raw samples/s of the signal-generation hot path on N B200s (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (libsqg.so)
    python bench.py --impl reference --gpus N ...             # the reference's own CPU gen_sig on host cores

A step = one pass of the whole hot path (dwell pass, scans, signal kernel) over one resident batch of
synthetic reads; `value` has inputs and outputs resident in HBM, `e2e` goes through the host-buffer
C-ABI call (sqg_submit/sqg_wait) with the H2D/D2H copies inside the timed region.
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[2] (per-GPU shard; the north_star target is quoted on this profile)
    "dna-r10-prom": dict(profile="dna-r10-prom", k=9, rna=False, meth=False, config="configs[2]"),
    # configs[1]
    "dna-r9-prom": dict(profile="dna-r9-prom", k=6, rna=False, meth=False, config="configs[1]"),
    # configs[3] shape: whole transcripts, ragged, reversed output, fixed dwell 31
    "rna004-prom": dict(profile="rna004-prom", k=9, rna=True, meth=False, config="configs[3]"),
    # configs[4]: CpG-methylation model path - base-5 ranks, the reference's 5^9-entry (15.6 MB) R10 CpG table,
    # reads with 'M' at methylated CpG sites (synthetic methylation frequency 0.7 per site)
    "dna-r10-prom-meth": dict(profile="dna-r10-prom", k=9, rna=False, meth=True, config="configs[4]"),
}


def load_model(name, k, meth):
    """The reference's own built-in table for the workload (oracle/dump_models.py wrote it at build time through the
    compiled, unmodified reference; plain data, nothing of oracle/ is executed here), else a random-init table of the
    same shape.  Returns (interleaved float32 table, description)."""
    p = os.path.join(ROOT, "oracle", "_ref", "models", name + ".f32")
    n = (5 if meth else 4) ** k
    if os.path.exists(p):
        m = np.fromfile(p, dtype=np.float32)
        if m.size == 2 * n:
            return m, "the reference's built-in table"
    return synth_model(n), "random-init table (reference tables not built here)"


def synth_model(num_kmer, seed=7):
    """random-init pore model of the right shape (level_mean ~ U(60,130) pA, level_stdv ~ U(1,4) pA)"""
    rs = np.random.RandomState(seed)
    m = np.empty(2 * num_kmer, dtype=np.float32)
    m[0::2] = rs.uniform(60, 130, num_kmer)
    m[1::2] = rs.uniform(1.0, 4.0, num_kmer)
    return m


def methylate(bases, off, freq, seed):
    """CpG sites of every read -> 'M' with probability freq (what methylate_dna(), src/genread.c:207-241, does to reads
    of a genome whose --meth-freq file gives that frequency at every site)"""
    rs = np.random.RandomState(seed)
    b = bases.copy()
    cg = np.nonzero((b[:-1] == ord("C")) & (b[1:] == ord("G")))[0]
    ends = off[1:] - 1   # a CG pair must not straddle two reads
    cg = cg[~np.isin(cg, ends)]
    b[cg[rs.random_sample(cg.size) < freq]] = ord("M")
    return b


def synth_reads(n_reads, mean_len, rna, seed, genome_mb=64, with_coords=False):
    """Reads as the reference's sampler would cut them from a synthetic i.i.d. ACGT genome: length ~
    Gamma(2, rlen/2) (src/sim.c:243), uniform position, clipped at the contig end, <200 nt rejected
    (src/genread.c:125-154), strand coin + reverse complement.  RNA: whole 'transcripts' of ragged length
    (283..6943 nt like the sequins, mean ~1355)."""
    rs = np.random.RandomState(seed)
    g = rs.randint(0, 4, genome_mb << 20).astype(np.uint8)
    comp = np.array([3, 2, 1, 0], dtype=np.uint8)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    if rna:
        lens = np.clip(rs.gamma(2.2, 1355 / 2.2, n_reads), 283, 6943).astype(np.int64)
    else:
        lens = rs.gamma(2.0, mean_len / 2.0, int(n_reads * 1.1) + 16).astype(np.int64)
        lens = lens[lens >= 200][:n_reads]
    pos = (rs.random_sample(len(lens)) * (len(g) - 1)).astype(np.int64)
    lens = np.minimum(lens, len(g) - pos)
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    bases = np.empty(off[-1], dtype=np.uint8)
    strand = rs.randint(0, 2, len(lens))
    for i in range(len(lens)):
        s = g[pos[i]:pos[i] + lens[i]]
        if not rna and strand[i]:
            s = comp[s[::-1]]
        bases[off[i]:off[i + 1]] = lut[s]
    if with_coords:  # the same reads as (contig, len, pos, strand) against the genome, for sqg_submit_coords
        from squigulator_b200.api import COORD_DTYPE
        co = np.zeros(len(lens), dtype=COORD_DTYPE)
        co["len"], co["pos"] = lens, pos
        co["strand"] = np.where((strand != 0) & (not rna), ord("-"), ord("+"))
        return bases, off, lut[g], co
    return bases, off


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region (NVML, 20 ms period)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz, self.power = index, False, [], set(), None, []

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            while not self.stop_flag:
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.02)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "power_w_max": max(self.power) if self.power else None, "samples": len(self.sm),
                "reasons": sorted(self.reasons)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own gen_sig (oracle/_ref/libsqref.so, built from /root/reference unmodified), else
# the plain-C oracle port.  The only place bench.py executes anything under oracle/.
def cpu_reference_run(wl, bases, off, seed, max_threads=None):
    from tests import helpers as H
    prof, flags = H.PRESETS[wl["profile"]]
    cores = os.cpu_count() or 1
    nthreads = min(cores, max_threads or cores)
    n_reads = len(off) - 1
    so = os.path.join(ROOT, "oracle", "_ref", "libsqref.so")
    if os.path.exists(so):
        lib = C.CDLL(so)
        lib.sqref_open.restype = C.c_void_p
        lib.sqref_open.argtypes = [C.POINTER(H.Profile), C.c_uint32, C.c_int64, C.c_int32, C.c_float, C.c_int,
                                   C.c_char_p, C.c_char_p, C.c_int]
        lib.sqref_run_batch.restype = C.c_int64
        lib.sqref_run_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
        lib.sqref_close.argtypes = [C.c_void_p]
        p = H.make_profile(prof)
        h = lib.sqref_open(C.byref(p), flags, seed, nthreads, 1.0, 0, None, None, 0)  # built-in tables of the reference
        lens = np.ascontiguousarray(np.diff(off).astype(np.int32))
        t0 = time.perf_counter()
        total = lib.sqref_run_batch(h, bases.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p),
                                    lens.ctypes.data_as(C.c_void_p), n_reads, nthreads)
        dt = time.perf_counter() - t0
        lib.sqref_close(h)
        return dict(kind="reference", cores=nthreads, samples=int(total), seconds=dt)
    # port: single-threaded oracle, legacy (reference-exact) RNG
    o = H.Oracle(H.load_oracle(), prof, flags, wl["k"], 4 ** wl["k"], synth_model(4 ** wl["k"]), seed, H.RNG_LEGACY)
    raw = bases.tobytes()
    t0 = time.perf_counter()
    total = 0
    for i in range(n_reads):
        total += len(o.gen_sig(raw[off[i]:off[i + 1]], read_index=i)["sig"])
    dt = time.perf_counter() - t0
    o.close()
    return dict(kind="port", cores=1, samples=total, seconds=dt)


def run_reference_arm(args, wl, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # ~2 s of CPU work per step so that a --steps 50 run still ends within a few minutes
    reads_per_step = args.cpu_reads if args.cpu_reads else max(cores * 200, 64)
    bases, off = synth_reads(reads_per_step, args.read_len, wl["rna"], seed=1234)
    times, samples = [], 0
    for i in range(args.warmup + args.steps):
        r = cpu_reference_run(wl, bases, off, seed=1)
        if i >= args.warmup:
            times.append(r["seconds"])
            samples += r["samples"]
    v = samples / sum(times)
    dm = dict(__import__("tests.helpers", fromlist=["PRESETS"]).PRESETS[wl["profile"]][0])["dwell_mean"]
    sample = f"{reads_per_step} reads (~{r['samples']} samples) per step, gen_sig only (no SLOW5 encode), {r['cores']} threads"
    line = {"impl": "reference", "metric": "int16_raw_samples_per_s", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "gbases_per_s": v / dm / 1e9,
            "config": {"workload": f"{wl['config']}: -x {wl['profile']}, reads cut from a synthetic iid genome, mean length {args.read_len}",
                       "reads_per_step": reads_per_step},
            "cpu_baseline": {"value": v, "unit": "samples/s", "cores": r["cores"], "kind": r["kind"], "sample": sample},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_json(line)


# ------------------------------------------------------------------------------------------------
def measure_gpu(args, wl, gen, sq, dist, rank, world, device, reads_per_step, steps, warmup, with_e2e=True):
    import torch
    from squigulator_b200.api import PROFILES
    prof = PROFILES[wl["profile"]][0]
    bases, off, genome, coords = synth_reads(reads_per_step, args.read_len, wl["rna"], seed=1000 + rank, with_coords=True)
    if wl["meth"]:
        bases = methylate(bases, off, 0.7, seed=2000 + rank)
    n_reads = len(off) - 1
    first = rank * n_reads  # read-index range of this GPU: the Philox counter makes the job independent of N
    out = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{device}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{device}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- kernel-only: resident batch ----
    db = gen.dev_batch(bases, off, first_read_index=first)
    gen.dev_batch_run(db, max(warmup, 3))
    info = gen.dev_batch_info(db)
    sampler = ClockSampler(device)
    l0 = gen.launch_count()
    barrier()
    sampler.start()
    w0 = time.perf_counter()
    ms_total, ms_k4 = gen.dev_batch_run(db, steps)
    barrier()
    wall = time.perf_counter() - w0
    sampler.stop_flag = True
    sampler.join()
    launches = gen.launch_count() - l0
    ms = max_over_ranks(ms_total)
    samples_all = sum_over_ranks(float(info["samples"]))
    out.update(value=samples_all * steps / (ms * 1e-3), ms_per_step=ms / steps, wall_s=wall, clocks=sampler.summary(),
               gpu_launches=int(launches), samples_per_step_per_gpu=info["samples"], kmers_per_step_per_gpu=info["kmers"],
               reads_per_step_per_gpu=n_reads)
    # roofline of the signal kernel on this rank: algorithmic bytes = 2 B/sample + 1 B per k-mer (one base)
    peak, how = measured_peak()
    alg_bytes = 2.0 * info["samples"] + 1.0 * info["kmers"]
    k4_ms = ms_k4 / steps
    # DRAM traffic of one launch from the committed ncu --set full capture of this very configuration, else null
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r02_signal_kernel_ncu.json")
    if os.path.exists(tp):
        t = json.load(open(tp))
        if t.get("workload") == wl["profile"] and t.get("reads_per_step") == n_reads and t.get("samples") == info["samples"]:
            traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
    out["roofline"] = {"bound": "hbm", "kernel": "signal_kernel", "achieved": alg_bytes / (k4_ms * 1e-3) / 1e9, "peak": peak,
                       "unit": "GB/s", "frac": alg_bytes / (k4_ms * 1e-3) / 1e9 / peak, "traffic": traffic,
                       "peak_source": how, "kernel_ms": k4_ms, "kernel_share_of_step": ms_k4 / ms_total,
                       "algorithmic_bytes_per_launch": alg_bytes}
    gen.dev_batch_destroy(db)

    # ---- end to end: host buffers through the asynchronous C-ABI (H2D + kernels + D2H every step) ----
    if with_e2e:
        e_reads = min(n_reads, args.e2e_reads)
        e_off = np.ascontiguousarray(off[:e_reads + 1])
        nb = int(e_off[-1])
        lib = sq.load_library()
        hp = lib.sqg_host_alloc(nb)  # pinned input buffer
        hb = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint8)), shape=(nb,))
        hb[:] = bases[:nb]
        e_steps = max(6, min(steps, 24))
        checks = 0

        def pipeline(n, want=0, by_coords=False):
            nonlocal checks
            inflight, tot, d2h = [], 0, 0

            def finish(t):
                nonlocal tot, d2h, checks
                r = gen.wait(t)
                tot += r.total_samples
                if want & 2:   # SQG_WANT_SVB: svb-zd streams instead of raw int16
                    d2h = int(r.svb_off[r.n_reads]) + r.n_reads * 44
                    checks ^= int(r.svb[0])
                else:
                    d2h = int(r.sig_off[r.n_reads - 1] + ((r.len_raw_signal[r.n_reads - 1] + 63) & ~63)) * 2 + r.n_reads * 28
                    checks ^= int(r.signal[0])  # touch the result on the host
                gen.release(t)

            for i in range(n):
                if by_coords:
                    inflight.append(gen.submit_coords(e_coords, first_read_index=first, want=want))
                else:
                    inflight.append(gen.submit(hb, e_off, first_read_index=first, want=want))
                if len(inflight) == 3:
                    finish(inflight.pop(0))
            for t in inflight:
                finish(t)
            return tot, d2h

        # the ceiling this job can reach: plain cudaMemcpyAsync device -> pinned host, all ranks at once, no kernels
        ceil_bytes = 1 << 30
        hsrc = torch.empty(ceil_bytes, dtype=torch.uint8, device=f"cuda:{device}")
        hdst = torch.empty(ceil_bytes, dtype=torch.uint8, pin_memory=True)
        hdst.copy_(hsrc, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(4):
            hdst.copy_(hsrc, non_blocking=True)
        barrier()
        d2h_gbs_rank = 4 * ceil_bytes / max_over_ranks(time.perf_counter() - t0) / 1e9   # per rank, while all ranks copy
        del hsrc, hdst
        out["d2h_ceiling"] = {"gbs_per_gpu": d2h_gbs_rank, "gbs_all": d2h_gbs_rank * world,
                              "how": "4 x 1 GiB cudaMemcpyAsync device -> pinned host on every rank at once (max over ranks)"}
        pipeline(3)
        barrier()
        t0 = time.perf_counter()
        tot, d2h = pipeline(e_steps)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        out["e2e"] = {"value": sum_over_ranks(float(tot)) / dt, "unit": "samples/s",
                      "h2d_bytes_per_step": nb + (e_reads * 64), "d2h_bytes_per_step": d2h,
                      "steps": e_steps, "reads_per_step_per_gpu": e_reads, "slots": 3,
                      "d2h_gbs_per_gpu": d2h * e_steps / dt / 1e9,
                      "frac_of_d2h_ceiling": d2h * e_steps / dt / 1e9 / d2h_gbs_rank,
                      "api": "sqg_submit/sqg_wait/sqg_release (pinned host bases in, pinned host int16 out)"}
        # the same job with SQG_WANT_SVB: the signal crosses PCIe as slow5lib's svb-zd stream
        pipeline(3, want=2)
        barrier()
        t0 = time.perf_counter()
        tot, d2h = pipeline(e_steps, want=2)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        out["e2e_svb"] = {"value": sum_over_ranks(float(tot)) / dt, "unit": "samples/s",
                          "h2d_bytes_per_step": nb + (e_reads * 64), "d2h_bytes_per_step": d2h, "steps": e_steps,
                          "frac_of_d2h_ceiling": d2h * e_steps / dt / 1e9 / d2h_gbs_rank,
                          "api": "same call with SQG_WANT_SVB: svb-zd streams (zig-zag delta + StreamVByte, bit-identical to slow5lib's) out"}
        # both ends shrunk: reads named by coordinates against the genome resident in HBM, svb-zd out
        gen.load_genome([genome.tobytes()])
        e_coords = np.ascontiguousarray(coords[:e_reads])
        tot0 = tot
        pipeline(3, want=2, by_coords=True)
        barrier()
        t0 = time.perf_counter()
        tot, d2h = pipeline(e_steps, want=2, by_coords=True)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        assert tot == tot0, "coordinate batches must generate exactly the reads of the host-bases batches"
        out["e2e_coords_svb"] = {"value": sum_over_ranks(float(tot)) / dt, "unit": "samples/s",
                                 "h2d_bytes_per_step": e_reads * (24 + 8), "d2h_bytes_per_step": d2h, "steps": e_steps,
                                 "api": "sqg_submit_coords with SQG_WANT_SVB: 24-byte coordinates in, reads cut out of the "
                                        "HBM-resident genome on the GPU, svb-zd streams out"}
        lib.sqg_host_free(hp)
    return out


_JSON_FD = None


def emit_json(line):
    """The ONE line of stdout.  Everything else this process (or NCCL, or a child) writes to fd 1 goes to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)   # keep the real stdout for the JSON line ...
    os.dup2(2, 1)          # ... and send every other write to fd 1 (e.g. NCCL's version banner) to stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="dna-r10-prom", choices=sorted(WORKLOADS))
    ap.add_argument("--reads-per-step", type=int, default=32768, help="reads in the resident batch of each GPU")
    ap.add_argument("--e2e-reads", type=int, default=8192)
    ap.add_argument("--read-len", type=int, default=10000)
    ap.add_argument("--cpu-reads", type=int, default=0)
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workloads and the CPU baseline")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]

    if args.impl == "reference":
        run_reference_arm(args, wl, rank)
        return

    import torch
    import torch.distributed as dist
    import squigulator_b200 as sq
    from squigulator_b200.api import PROFILES

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path in squigulator_b200)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries the one JSON line only
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    # NUMA: bind this rank's threads to the CPUs next to its GPU before any pinned buffer is touched (first touch decides
    # where the pages live, and the D2H direction is what bounds the end-to-end figure)
    try:
        import pynvml as nv
        nv.nvmlInit()
        nv.nvmlDeviceSetCpuAffinity(nv.nvmlDeviceGetHandleByIndex(local))
        affinity = f"nvml ideal CPUs ({len(os.sched_getaffinity(0))} of {os.cpu_count()})"
    except Exception as e:  # pragma: no cover
        affinity = f"not set ({type(e).__name__})"

    table_src = {}

    def make_gen(w, rng_mode=0, name=None):
        # pore-model table: loaded on rank 0, broadcast once over NCCL, handed to the library as a device pointer
        name = name or [k for k, v in WORKLOADS.items() if v is w][0]
        n = (5 if w["meth"] else 4) ** w["k"]
        t = torch.empty(2 * n, dtype=torch.float32, device=f"cuda:{local}")
        if rank == 0:
            m, table_src[name] = load_model(name, w["k"], w["meth"])
            t.copy_(torch.from_numpy(m))
        if world > 1:
            dist.broadcast(t, src=0)
        torch.cuda.synchronize()
        prof, flags = PROFILES[w["profile"]]
        return sq.SignalGenerator(dict(prof), None, w["k"], flags=flags, seed=1, device=local, n_slots=3, meth=w["meth"],
                                  rng_mode=rng_mode, device_model_ptr=t.data_ptr()), t

    gen, _keep = make_gen(wl)
    res = measure_gpu(args, wl, gen, sq, dist, rank, world, local, args.reads_per_step, args.steps, args.warmup)
    # HBM write ceiling (store-only kernel over 8 GiB) for context
    store_ms = gen.bench_store(8 << 30, 5)
    store_ms = gen.bench_store(8 << 30, 10)
    res["store_only_gbs"] = (8 << 30) * 10 / (store_ms * 1e-3) / 1e9

    # ---- N GPUs produce exactly the one-GPU output: every rank generates its shard of ONE common read list, rank 0 also
    # the whole list, per-read digests are compared (outside every timed region) ----
    shard_equal = None
    if world > 1:
        import hashlib
        from squigulator_b200.shard import shard_range
        cb, co = synth_reads(64 * world, 3000, wl["rna"], seed=4242, genome_mb=8)
        if wl["meth"]:
            cb = methylate(cb, co, 0.7, seed=4243)
        nr = len(co) - 1
        lo, hi = shard_range(nr, rank, world)
        raw = cb.tobytes()
        mine = gen.gen_batch([raw[co[i]:co[i + 1]] for i in range(lo, hi)], first_read_index=lo)
        dig = [(lo + i, hashlib.sha256(r["sig"].tobytes()).hexdigest(), r["offset"]) for i, r in enumerate(mine)]
        allv = [None] * world
        dist.all_gather_object(allv, dig)
        if rank == 0:
            whole = gen.gen_batch([raw[co[i]:co[i + 1]] for i in range(nr)], first_read_index=0)
            exp = {i: (hashlib.sha256(r["sig"].tobytes()).hexdigest(), r["offset"]) for i, r in enumerate(whole)}
            got = {i: (h, o) for part in allv for (i, h, o) in part}
            shard_equal = len(got) == nr and all(got[i] == exp[i] for i in range(nr))
    gen.close()

    extra = {}
    if not args.no_extra:
        for name in sorted(WORKLOADS):
            if name == args.workload:
                continue
            w2 = WORKLOADS[name]
            g2, _k2 = make_gen(w2)
            rps = args.reads_per_step if not w2["rna"] else args.reads_per_step * 4
            r2 = measure_gpu(args, w2, g2, sq, dist, rank, world, local, rps, max(10, args.steps // 2), 3, with_e2e=False)
            g2.close()
            extra[name] = {"config": w2["config"], "value": r2["value"], "ms_per_step": r2["ms_per_step"],
                           "roofline_frac": r2["roofline"]["frac"], "kernel_ms": r2["roofline"]["kernel_ms"],
                           "samples_per_step_per_gpu": r2["samples_per_step_per_gpu"], "table": table_src.get(name)}
        # SQG_RNG_LEGACY (the reference's own minstd streams by jump-ahead: the parity instrument, not the fast path),
        # whole path, same metric - so that its cost is on record
        wl9 = WORKLOADS["dna-r9-prom"]
        g3, _k3 = make_gen(wl9, rng_mode=1, name="dna-r9-prom")
        lb, lo_ = synth_reads(512, args.read_len, False, seed=77 + rank, genome_mb=8)
        raw = lb.tobytes()
        reads = [raw[lo_[i]:lo_[i + 1]] for i in range(len(lo_) - 1)]
        g3.gen_batch(reads[:32], first_read_index=0)
        barrier_t0 = time.perf_counter()
        rr = g3.gen_batch(reads, first_read_index=0)
        dt = time.perf_counter() - barrier_t0
        g3.close()
        extra["legacy-rng dna-r9-prom"] = {"config": "configs[1] profile, SQG_RNG_LEGACY, host buffers in and out (one synchronous batch of 512 reads)",
                                           "value": sum(len(r["sig"]) for r in rr) / dt, "unit": "samples/s per GPU"}
        # the per-read seam: sqg_gen_sig(), the call with gen_sig()'s own shape (src/gensig.c:346) for a host that cannot
        # batch - one read per call, planning + signal kernels + copies + a malloc'd result each time
        g4, _k4 = make_gen(wl)
        g4.gen_sig(reads[0], read_index=0)
        t0 = time.perf_counter()
        ns = sum(len(g4.gen_sig(r, read_index=i)["sig"]) for i, r in enumerate(reads[:128]))
        dt = time.perf_counter() - t0
        g4.close()
        extra["per-read seam sqg_gen_sig"] = {"config": f"{wl['config']} profile, one synchronous call per read (128 reads of ~{args.read_len} nt)",
                                              "value": ns / dt, "unit": "samples/s per GPU", "calls_per_s": 128 / dt}

    cpu = None
    if rank == 0 and world == 1 and not args.no_extra:
        cores = os.cpu_count() or 1
        n_cpu = args.cpu_reads if args.cpu_reads else max(cores * 1400, 64)  # ~10-20 s of gen_sig on all cores
        cb, co = synth_reads(n_cpu, args.read_len, wl["rna"], seed=1234)
        r = cpu_reference_run(wl, cb, co, seed=1)
        cpu = {"value": r["samples"] / r["seconds"], "unit": "samples/s", "cores": r["cores"], "kind": r["kind"],
               "sample": f"{n_cpu} reads of the same workload ({r['samples']} samples, {r['seconds']:.1f} s), gen_sig only, {r['cores']} host threads"}

    if rank == 0:
        prof = PROFILES[wl["profile"]][0]
        line = {"metric": "int16_raw_samples_per_s", "value": res["value"], "unit": "samples/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "gbases_per_s": res["value"] / prof["dwell_mean"] / 1e9,
                "config": {"workload": f"{wl['config']} per-GPU shard: -x {wl['profile']}, reads cut (gamma length, mean {args.read_len}) "
                                       f"from a synthetic iid ACGT genome, {wl['k']}-mer table; step = one resident batch",
                           "reads_per_step_per_gpu": res["reads_per_step_per_gpu"],
                           "samples_per_step_per_gpu": res["samples_per_step_per_gpu"],
                           "l2": "output per step (GBs) far exceeds the 126 MB L2; no flush needed",
                           "table": table_src.get(args.workload), "cpu_affinity": affinity,
                           "rng": "philox4x32-7", "parallelism": f"reads sharded over {world} GPU(s), no hot-path collective"},
                "clocks": res["clocks"], "e2e": res.get("e2e"), "e2e_svb": res.get("e2e_svb"), "e2e_coords_svb": res.get("e2e_coords_svb"), "gpu_launches": res["gpu_launches"],
                "roofline": res["roofline"], "cpu_baseline": cpu, "store_only_gbs": res["store_only_gbs"],
                "d2h_ceiling": res.get("d2h_ceiling"), "shard_equal": shard_equal,
                "wall_s_timed_region": res["wall_s"], "other_workloads": extra}
        emit_json(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
