set -u
O=gpurun_out/r02k; mkdir -p $O
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -5 $O/bench.err
python - <<PY
import json
d=json.load(open("$O/bench.json"))
print("value", d["value"], "roofline", d["roofline"]["frac"], d["roofline"]["kernel"], "e2e", d["e2e"]["value"] if d["e2e"] else None)
print("parity", d["parity_check"])
for k,v in (d["configs"] or {}).items():
    print(k, v["value"], v["ms"], v["roofline"]["frac"], v["roofline"]["kernel"], v["parity_check"]["ok"] if v["parity_check"] else None, v.get("strong_scaling"))
print("cpu", d["cpu_baseline"]["value"] if d["cpu_baseline"] else None, "numa", d["numa"])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; cut -c1-400 $O/bench_ref.json
grep -c fqtk_b200 /proc/self/maps > /dev/null
