B="python bench.py --no-cpu-baseline --no-e2e --no-brute --no-routing --no-configs --no-parity-check"
$B --steps 3 --warmup 3 2>/tmp/err.txt | python -c "
import json,sys;d=json.loads(sys.stdin.read());w=d['bgzf']['whole_data_path'];print(w['ms'],w['mreads_per_s'],w['one_call'])"; tail -3 /tmp/err.txt
