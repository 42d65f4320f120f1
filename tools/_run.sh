mkdir -p gpurun_out/r02z
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_bgzf.py -m gpu -x -q -k "tiny or one_byte or fastq_3_blocks" > gpurun_out/r02z/sanitize.log 2>&1; tail -25 gpurun_out/r02z/sanitize.log
timeout 900 python -m pytest tests/test_bgzf.py -m gpu -x -q 2>&1 | tail -25
