#!/bin/bash
# (the .ncu-rep files stay on the box under /tmp: gpurun brings back at most 64 MiB; the summaries carry what is read here)
# One gpurun session: GPU tests, bench, ncu launch list + full captures of the dominant kernels.  Outputs -> gpurun_out/$TAG.
# usage: tools/gpu_session.sh [tag] [steps...]   steps default: tests bench launches ncu
set -u
TAG=${1:-r02}; shift || true
STEPS=${*:-tests bench launches ncu}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > "$OUT/gpu.csv" 2>&1
B="python bench.py --no-cpu-baseline --no-e2e --no-brute --no-routing --no-bgzf --no-configs --no-parity-check"
cap() {  # name, kernel regex, bench args...
  name=$1; regex=$2; shift 2
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$regex -s ${SKIP:-2} -c 1 -o "/tmp/prof_$name" -f $B --steps 2 --warmup 1 "$@" > "$OUT/ncu_$name.log" 2>&1
  echo "ncu $name rc=$?"
  python tools/ncu_summary.py "/tmp/prof_$name.ncu-rep" --sass --min 0.3 > "$OUT/summary_$name.txt" 2>&1
}
for s in $STEPS; do
  echo "=== $s ($(date +%T)) ==="
  case $s in
    smoke) timeout 600 python __graft_entry__.py --smoke > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?"; tail -3 "$OUT/smoke.log";;
    tests) timeout 1800 python -m pytest tests -m gpu -q --maxfail=8 -x --durations=8 > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?"; tail -14 "$OUT/pytest_gpu.log";;
    bench) timeout 900 python bench.py --steps 10 --warmup 3 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?"; tail -3 "$OUT/bench.err"; cut -c1-600 "$OUT/bench.json";;
    benchref) timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"; echo "benchref rc=$?"; cut -c1-300 "$OUT/bench_ref.json";;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/launches.csv" \
        python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-reads 8388608 > "$OUT/ncu_launches_bench.log" 2>&1; echo "ncu launches rc=$?";;
    ncu)
      cap cfg3_k_probe3 k_probe3
      cap cfg2_k_probe3 k_probe3 --config 2
      cap cfg4_k_probe5 k_probe5 --config 4 --reads 268435456
      cap cfg5_k_probe4 k_probe4 --config 5 --reads 268435456
      cap cfg3_k_brute_sliced k_brute_sliced --mode brute --reads 67108864
      B_SAVE=$B
      B="python bench.py --no-cpu-baseline --no-e2e --no-brute --no-bgzf --no-configs --no-parity-check"
      cap cfg3_k_route_scatter k_route_scatter_tile2
      B="python bench.py --no-cpu-baseline --no-e2e --no-brute --no-routing --no-configs --no-parity-check"
      cap fastq_k_bgzf_deflate k_bgzf_deflate
      B=$B_SAVE
      ls -la "$OUT";;
  esac
done
