set -u
O=gpurun_out/r02j; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "kernel_variants or configs_reduced_n or fuzz or out_of_alphabet or full_size_properties" 2>&1 | tail -30 > $O/pytest.txt; tail -5 $O/pytest.txt
run() { # name cfg reads env...
  name=$1; cfg=$2; reads=$3; shift 3
  env "$@" timeout 600 python bench.py --config $cfg --reads $reads --steps 5 --warmup 2 --no-cpu-baseline --no-e2e --no-brute > $O/$name.json 2> $O/$name.err
  python -c "
import json
try:
    d=json.load(open('$O/$name.json')); print('$name', d['roofline']['kernel'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['memo_table']['l2_table_bytes'])
except Exception as e: print('$name failed', e)"
}
for f in 3 19 51; do run cfg5_load50_f$f 5 1000000000 FQTK_B200_G4_LOAD=50 FQTK_B200_G4_FLAGS=$f; done
run cfg5_load60_f3 5 1000000000 FQTK_B200_G4_LOAD=60 FQTK_B200_G4_FLAGS=3
run cfg4_load40_f3 4 1000000000 FQTK_B200_G4_LOAD=40 FQTK_B200_G4_FLAGS=3
python bench.py --config 4 --cuckoo 0 --steps 5 --warmup 2 --no-cpu-baseline --no-e2e --no-brute 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg4 k_probe2', d['roofline']['kernel_ms'], d['roofline']['frac'])"
