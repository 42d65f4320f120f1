set -u
timeout 900 python -m pytest tests/test_bgzf.py tests/test_cpp_mirror.py -m gpu -x -q 2>&1 | tail -3
B="python bench.py --no-cpu-baseline --no-e2e --no-brute --no-routing --no-configs --no-parity-check"
for L in 2 3; do
FQTK_B200_BENCH_LANES=$L $B --steps 3 --warmup 3 2>/tmp/err.txt | python -c "
import json,sys;d=json.loads(sys.stdin.read());b=d['bgzf'];w=b['whole_data_path'];print($L, 'one', w['one_call']['ms'], 'lanes', w['two_lanes']['ms_per_batch'], 'dev', b['device']['ms'], 'host_call', b['host_call']['ms'])"; tail -2 /tmp/err.txt
done
FQTK_B200_TRACE=1 $B --steps 3 --warmup 3 2>&1 >/dev/null | grep demux_chunks | head -12
