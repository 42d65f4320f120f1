set -u
python -m pytest tests -m gpu -x -q -k "kats or fuzz or configs_reduced_n or long_barcodes or many_samples or length_rules or single_sample or extreme" 2>&1 | tail -5
for c in 3 4 5 2; do
python bench.py --config $c --mode brute --reads 67108864 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --no-brute --no-configs --no-parity-check 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg$c sliced', d['roofline']['kernel'], d['roofline']['kernel_ms'], d['roofline']['pair_compares_per_s'])"
FQTK_B200_BRUTE_V1=1 python bench.py --config $c --mode brute --reads 67108864 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --no-brute --no-configs --no-parity-check 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg$c v1    ', d['roofline']['kernel'], d['roofline']['kernel_ms'], d['roofline']['pair_compares_per_s'])"
done
