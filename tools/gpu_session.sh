#!/bin/bash
# One gpurun session: smoke, microbench, GPU tests, bench, ncu launch list + full capture.  Outputs -> gpurun_out/.
# usage: tools/gpu_session.sh [tag] [steps...]   steps default: smoke micro tests bench ncu
set -u
TAG=${1:-r01}; shift || true
STEPS=${*:-smoke micro tests bench ncu}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > "$OUT/gpu.csv" 2>&1
nproc > "$OUT/nproc.txt"; free -g >> "$OUT/nproc.txt"; lscpu | head -20 >> "$OUT/nproc.txt"
for s in $STEPS; do
  echo "=== $s ($(date +%T)) ==="
  case $s in
    smoke) timeout 600 python __graft_entry__.py --smoke > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?"; tail -3 "$OUT/smoke.log";;
    micro) timeout 120 tools/microbench > "$OUT/microbench.txt" 2>&1; echo "micro rc=$?"; cat "$OUT/microbench.txt";;
    tests) timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 -x --durations=12 > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?"; tail -40 "$OUT/pytest_gpu.log";;
    bench) timeout 900 python bench.py --steps 10 --warmup 3 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?"; cat "$OUT/bench.json"; tail -5 "$OUT/bench.err";;
    ab) for c in 2 3 0; do timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-brute --cuckoo $c > "$OUT/bench_ab_cuckoo$c.json" 2> "$OUT/bench_ab_cuckoo$c.err"; echo "ab cuckoo=$c rc=$?"; python -c "import json,sys; d=json.load(open('$OUT/bench_ab_cuckoo$c.json')); print(d['roofline']['kernel'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['value'])"; done;;
    benchref) timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"; echo "benchref rc=$?"; cat "$OUT/bench_ref.json";;
    bench2) for c in 2 4 5; do timeout 900 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > "$OUT/bench_cfg$c.json" 2> "$OUT/bench_cfg$c.err"; echo "bench cfg$c rc=$?"; cat "$OUT/bench_cfg$c.json"; done;;
    ncu)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" \
        python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-reads 8388608 > "$OUT/ncu_launches_bench.log" 2>&1; echo "ncu launches rc=$?"
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_probe -s 1 -c 1 -o "$OUT/prof_probe" -f \
        python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-brute > "$OUT/ncu_probe.log" 2>&1; echo "ncu probe rc=$?"
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_brute -s 1 -c 1 -o "$OUT/prof_brute" -f \
        python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --mode brute --reads 67108864 > "$OUT/ncu_brute.log" 2>&1; echo "ncu brute rc=$?"
      ls -la "$OUT";;
  esac
done
