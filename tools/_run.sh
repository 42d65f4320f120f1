python -m pytest tests -m gpu -x -q -k "fastq or batched_pipeline" 2>&1 | tail -8
python bench.py --steps 3 --warmup 3 --no-configs --no-cpu-baseline --no-brute 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps(d['fastq_ingest'], indent=1)); print(d['e2e']['value'])"
