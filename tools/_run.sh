B="python bench.py --no-cpu-baseline --no-brute --no-routing --no-bgzf --no-configs --no-parity-check --e2e-reads 8388608"
$B --steps 3 --warmup 3 2>/tmp/err.txt | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(json.dumps(d['fastq_ingest'],indent=1))"; tail -3 /tmp/err.txt
