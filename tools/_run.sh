mkdir -p gpurun_out/r02p
true
B="python bench.py --no-cpu-baseline --no-e2e --no-brute --no-configs --no-parity-check"
$B --steps 5 --warmup 3 > gpurun_out/r02p/bench.json 2> gpurun_out/r02p/bench.err; python -c "
import json;d=json.load(open('gpurun_out/r02p/bench.json'));print(d['routing'])"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_route --csv --log-file gpurun_out/r02p/launches_route.csv $B --steps 1 --warmup 1 > /dev/null 2>&1
grep -c k_route gpurun_out/r02p/launches_route.csv; python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02p/launches_route.csv')) if len(r)>5 and 'k_route' in ''.join(r)]
for r in rows[-8:]: print(r[4][:40], r[-1])
PY
