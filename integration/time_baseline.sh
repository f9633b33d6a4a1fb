#!/bin/bash
# BASELINE.md section 4, items 2-3, on this machine: the unmodified reference (oracle/_ref/squigulator) and the same CLI
# with process_db() on the GPU (oracle/_ref/squigulator_sqg, SQG_GPU=1), timed by the binary's own "[main] Real time" line.
#   integration/time_baseline.sh [outdir]
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="$ROOT/oracle/_ref"
OUT="$(mkdir -p "${1:-$ROOT/gpurun_out}" && cd "${1:-$ROOT/gpurun_out}" && pwd)"
cd "$REF"
NPROC=$(nproc)
samples() { python3 - "$1" <<'PY'
import sys
n = 0
for l in open(sys.argv[1], "rb"):
    if l[:1] in b"#@": continue
    n += int(l.split(b"\t")[6])
print(n)
PY
}
run() {  # label, env-prefix, binary, args...
    local label="$1" envp="$2" bin="$3"; shift 3
    local t0=$(date +%s%N)
    env $envp "$bin" "$@" > "$OUT/tb_$label.log" 2>&1
    local rc=$? t1=$(date +%s%N)
    local real=$(grep -o "Real time: [0-9.]*" "$OUT/tb_$label.log" | awk '{print $3}')
    echo "$label rc=$rc wall_ms=$(( (t1 - t0) / 1000000 )) real_time_line=${real:-NA}"
}
echo "host: $NPROC cores, $(grep -m1 'model name' /proc/cpuinfo | cut -d: -f2)"
# item 2: config 1 verbatim (SLOW5 ASCII so that the samples can be counted from column 7; blow5 timing next to it)
run cfg1_cpu_slow5 "X=1" "$REF/squigulator" test/nCoV-2019.reference.fasta -x dna-r9-prom -n 100 --seed 1 -t 1 -o "$OUT/tb_cfg1_cpu.slow5"
run cfg1_cpu_blow5 "X=1" "$REF/squigulator" test/nCoV-2019.reference.fasta -x dna-r9-prom -n 100 --seed 1 -t 1 -o "$OUT/tb_cfg1_cpu.blow5"
run cfg1_gpu_slow5 "SQG_GPU=1" "$REF/squigulator_sqg" test/nCoV-2019.reference.fasta -x dna-r9-prom -n 100 --seed 1 -t 1 -o "$OUT/tb_cfg1_gpu.slow5"
run cfg1_gpu_blow5 "SQG_GPU=1" "$REF/squigulator_sqg" test/nCoV-2019.reference.fasta -x dna-r9-prom -n 100 --seed 1 -t 1 -o "$OUT/tb_cfg1_gpu.blow5"
echo "cfg1 samples: cpu $(samples "$OUT/tb_cfg1_cpu.slow5") gpu $(samples "$OUT/tb_cfg1_gpu.slow5")"
# item 3: all cores, end to end incl. BLOW5 (zlib + svb-zd) encode: a 30 Mb synthetic genome, R10, 20000 reads of ~10 kb
python3 - "$OUT/tb_genome.fa" <<'PY'
import sys, numpy as np
rs = np.random.RandomState(42)
with open(sys.argv[1], "w") as f:
    for c in range(3):
        f.write(f">chr{c+1}\n")
        s = np.frombuffer(b"ACGT", dtype=np.uint8)[rs.randint(0, 4, 10_000_000)].tobytes().decode()
        f.write(s + "\n")
PY
run all_cpu "X=1" "$REF/squigulator" "$OUT/tb_genome.fa" -x dna-r10-prom -n 20000 -r 10000 --seed 1 -t "$NPROC" -K 4096 -o "$OUT/tb_all_cpu.blow5"
run all_gpu "SQG_GPU=1" "$REF/squigulator_sqg" "$OUT/tb_genome.fa" -x dna-r10-prom -n 20000 -r 10000 --seed 1 -t "$NPROC" -K 4096 -o "$OUT/tb_all_gpu.blow5"
grep -h "libsqg stages" "$OUT/tb_all_gpu.log" | sed 's/\x1b\[[0-9;]*m//g'
# the same with BLOW5 records left uncompressed (signal still svb-zd): zlib out of the way
run all_gpu_nozlib "SQG_GPU=1 SQG_RECORD_PRESS=none" "$REF/squigulator_sqg" "$OUT/tb_genome.fa" -x dna-r10-prom -n 20000 -r 10000 --seed 1 -t "$NPROC" -K 4096 -o "$OUT/tb_all_gpu_nozlib.blow5"
grep -h "libsqg stages" "$OUT/tb_all_gpu_nozlib.log" | sed 's/\x1b\[[0-9;]*m//g'
# and read back by the unmodified slow5lib: record count and total samples of both files
for f in tb_all_gpu tb_all_gpu_nozlib; do python3 "$ROOT/integration/blow5_stat.py" "$OUT/$f.blow5" || true; done
ls -la "$OUT"/tb_all_*.blow5 | awk '{print $5, $9}'
rm -f "$OUT"/tb_genome.fa "$OUT"/tb_all_*.blow5 "$OUT"/tb_cfg1_*.blow5 "$OUT"/tb_cfg1_*.slow5
