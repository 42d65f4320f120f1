timeout 1500 python -m pytest tests/test_bgzf.py tests/test_gpu_parity.py -m gpu -x -q -k "whole_data_path or assign_fastq_chunks" 2>&1 | tail -30
