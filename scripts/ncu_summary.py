#!/usr/bin/env python3
"""Summarise an ncu capture of the signal kernel into profiles/ (run here, on the .ncu-rep brought back by gpurun).

    scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_signal_kernel_ncu --workload dna-r10-prom --reads 32768 --samples N
"""
import argparse, csv, io, json, subprocess, collections

ap = argparse.ArgumentParser()
ap.add_argument("rep"); ap.add_argument("out")
ap.add_argument("--workload", default="dna-r10-prom"); ap.add_argument("--reads", type=int, default=32768)
ap.add_argument("--samples", type=int, default=0)
a = ap.parse_args()

raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
def num(k):
    v, u = m[k]
    x = float(v.replace(",", ""))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}.get(u, 1)
    return x * scale
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__icc_request_hit_rate.pct",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active"]
stalls = {h: m[h][0] for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h}
js = {"kernel": m.get("Kernel Name", ("signal_kernel", ""))[0], "workload": a.workload, "reads_per_step": a.reads, "samples": a.samples,
      "duration_s": num("gpu__time_duration.sum"), "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
      "metrics": {k: " ".join(m[k]) for k in keep if k in m}, "stall_samples": stalls,
      "note": "one launch under ncu --set full --clock-control none (cold cache, serialised): shares and bytes, not bench timings"}
json.dump(js, open(a.out + ".json", "w"), indent=1)

src = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2, data = rows[1], rows[2:]
ci, si, so = h2.index("Instructions Executed"), h2.index("# Samples"), h2.index("Source")
tot = sum(int(r[ci]) for r in data); ts = sum(int(r[si]) for r in data)
ops = collections.Counter()
for r in data:
    op = r[so].split()[0] if not r[so].strip().startswith("@") else r[so].split()[1]
    ops[op.split(".")[0]] += int(r[ci])
with open(a.out + ".txt", "w") as f:
    f.write(f"# {js['kernel']}\n# workload {a.workload}, {a.reads} reads/step, {a.samples} samples per launch\n")
    f.write(f"duration {js['duration_s']*1e3:.3f} ms (under ncu), DRAM read {js['dram_bytes_read']/1e9:.3f} GB, write {js['dram_bytes_write']/1e9:.3f} GB\n")
    for k in keep:
        if k in m: f.write(f"{k} = {' '.join(m[k])}\n")
    f.write(f"\nwarp-instructions executed {tot} ({tot*32/max(a.samples,1):.1f} lane-instructions per sample), {len(data)} SASS instructions\n")
    f.write("opcode mix (executed warp-instructions): " + ", ".join(f"{k} {100*v/tot:.1f}%" for k, v in ops.most_common(14)) + "\n")
    f.write("\nstall samples:\n")
    for k, v in sorted(stalls.items(), key=lambda kv: -int(kv[1].replace(',', ''))):
        f.write(f"  {k.replace('smsp__pcsamp_warps_issue_stalled_', ''):24s} {v}\n")
    f.write("\ntop SASS instructions by stall samples:\n")
    for r in sorted(data, key=lambda r: -int(r[si]))[:25]:
        f.write(f"  {int(r[si]):7d} samples {int(r[ci]):11d} exec  {r[so].strip()[:90]}\n")
    hot = [x for x in ("UBLKCP", "SYNCS", "STG.E.EF.128", "LDS.U16", "IMAD.WIDE.U32", "F2I.TRUNC", "PRMT", "HADD2.F32") if any(x in r[so] for r in data)]
    f.write("\nSASS evidence present: " + ", ".join(hot) + "\n")
print("wrote", a.out + ".json", a.out + ".txt")
