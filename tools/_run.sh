timeout 900 python -m pytest tests/test_bgzf.py -m gpu -x -q 2>&1 | tail -3
B="python bench.py --no-cpu-baseline --no-e2e --no-brute --no-routing --no-configs --no-parity-check"
$B --steps 3 --warmup 3 2>/tmp/err.txt | python -c "
import json,sys;d=json.loads(sys.stdin.read());b=d['bgzf'];print(b['device'], b['host_call']['gb_per_s_in'], b['whole_data_path']['one_call'])"; tail -3 /tmp/err.txt
