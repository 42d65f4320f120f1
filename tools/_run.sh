timeout 1500 python -m pytest tests/test_bgzf.py -m gpu -x -q -k "fuzz" 2>&1 | tail -12
