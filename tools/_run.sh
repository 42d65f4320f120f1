timeout 1500 python -m pytest tests/test_bgzf.py tests/test_gpu_parity.py -m gpu -x -q -k "whole_data_path or header_rewrite or fastq" 2>&1 | tail -5
python tools/time_gpu_demux.py 2>&1 | tail -10
