set -u
B="python bench.py --no-cpu-baseline --no-e2e --no-brute --no-routing --no-configs --no-parity-check"
for L in 2 3 4; do
FQTK_B200_BENCH_LANES=$L $B --steps 3 --warmup 3 2>/tmp/err.txt | python -c "
import json,sys;d=json.loads(sys.stdin.read());b=d['bgzf'];w=b['whole_data_path'];print($L, w['one_call']['ms'], w['two_lanes']['ms_per_batch'])"; tail -2 /tmp/err.txt
done
