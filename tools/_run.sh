OUT=gpurun_out/r02san; mkdir -p $OUT
timeout 900 python -m pytest tests/test_bgzf.py -m gpu -x -q 2>&1 | tail -2
B="python bench.py --no-cpu-baseline --no-e2e --no-brute --no-routing --no-configs --no-parity-check"
$B --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());b=d['bgzf'];print(b['device'], b['host_call']['gb_per_s_in'])"
timeout 1200 compute-sanitizer --tool racecheck --log-file $OUT/bgzf_racecheck.log python -m pytest tests/test_bgzf.py -m gpu -x -q -k "fastq_3_blocks or long_runs or random_incompressible or short_tail or two_symbols" > $OUT/bgzf_racecheck.out 2>&1
echo "bgzf racecheck rc=$?"; tail -2 $OUT/bgzf_racecheck.out; tail -2 $OUT/bgzf_racecheck.log
