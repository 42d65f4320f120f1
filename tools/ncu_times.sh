#!/bin/bash
# usage: tools/ncu_times.sh <kernel-regex> <count> <out.csv> -- <command...>   (per-launch duration + DRAM bytes)
REGEX=$1; COUNT=$2; OUT=$3; shift 4
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:$REGEX -c $COUNT --csv --log-file $OUT "$@" > /dev/null 2>&1
python - "$OUT" <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
iK, iM, iV, iU, iID = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
agg = {}
for r in rows[1:]:
    agg.setdefault((r[iID], r[iK][:60]), {})[r[iM]] = (r[iV], r[iU])
for (i, k), m in agg.items():
    print(i, k, {a: " ".join(b) for a, b in m.items()})
PY
