FQTK_B200_ROUTE_T=1024 python -m pytest tests -m gpu -x -q -k "route_kernel_versions or route_is" 2>&1 | tail -2
B="python bench.py --no-cpu-baseline --no-e2e --no-brute --no-bgzf --no-configs --no-parity-check"
for t in 1024 512; do
FQTK_B200_ROUTE_T=$t $B --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print($t, d['routing']['ms'], d['routing']['roofline_frac'])"
done
