mkdir -p gpurun_out/r02p
python -m pytest tests -m gpu -x -q -k "route" 2>&1 | tail -2
B="python bench.py --no-cpu-baseline --no-e2e --no-brute --no-configs --no-parity-check"
$B --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['routing']['ms'], d['routing']['roofline_frac'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_route_scatter_tile2 -s 1 -c 1 -o /tmp/prof_route -f $B --steps 1 --warmup 1 > gpurun_out/r02p/ncu_route.log 2>&1
python tools/ncu_summary.py /tmp/prof_route.ncu-rep --sass --min 0.3 > gpurun_out/r02p/summary_route.txt 2>&1
