set -u
OUT=gpurun_out/s3d; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "brute_sliced_ties" 2>&1 | tail -4
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --kernel-regex kns=k_brute_sliced --log-file $OUT/brute_$tool.log python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "brute_sliced_ties and (8-33 or 16-3073 or 20-2049 or 32-1537)" > $OUT/brute_$tool.out 2>&1
  echo "brute $tool rc=$?"; tail -2 $OUT/brute_$tool.out; tail -2 $OUT/brute_$tool.log
  timeout 600 compute-sanitizer --tool $tool --kernel-regex kns=k_emit --log-file $OUT/emit_$tool.log python -m pytest tests/test_bgzf.py -m gpu -x -q -k "header_rewrite or demux_outputs_as_bgzf or two_lanes" > $OUT/emit_$tool.out 2>&1
  echo "emit $tool rc=$?"; tail -2 $OUT/emit_$tool.out; tail -2 $OUT/emit_$tool.log
done
