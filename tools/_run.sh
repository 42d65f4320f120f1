set -u
O=gpurun_out/r02l; mkdir -p $O
timeout 900 python bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline --no-brute > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -5 $O/bench.err
python -c "
import json
d=json.load(open('$O/bench.json')); print(json.dumps(d['e2e'], indent=1))"
timeout 600 python bench.py --single-process --gpus 1 --steps 5 --warmup 2 > $O/bench_sp.json 2> $O/bench_sp.err; echo "sp rc=$?"; tail -3 $O/bench_sp.err; cut -c1-1500 $O/bench_sp.json
tools/h2d_ceiling 1 134217728 16 4; tools/h2d_ceiling 1 134217728 8 2
