timeout 1500 python -m pytest tests/test_bgzf.py -m gpu -x -q -k "whole_data_path" 2>&1 | tail -40
