set -u
mkdir -p gpurun_out/s3b
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
B="python bench.py --no-cpu-baseline --no-e2e --no-routing --no-bgzf --no-configs --no-parity-check --no-brute --mode brute --steps 3 --warmup 2"
for c in 3 4 5 2; do
  $B --config $c --reads 67108864 2>gpurun_out/s3b/err_$c.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cfg', $c, 'ms', d['ms_per_step'], 'pairs/s', d['roofline'].get('pair_compares_per_s'), d['roofline'].get('kernel'))"
  tail -2 gpurun_out/s3b/err_$c.txt
done
