#!/bin/bash
# usage: tools/diag_units_classes.sh [n]   per-size-class durations of k_units_mma (ncu launch list) for the diagnostic builds
n=${1:-300000}
for f in "" "-DMMA_DIAG_NOSTORE" "-DMMA_DIAG_NOMMA" "-DMMA_DIAG_NOSTORE -DMMA_DIAG_NOMMA" "-DMMA_DIAG_NOSTORE -DMMA_DIAG_NOMMA -DMMA_DIAG_NOSCAL" "-DMMA_DIAG_BHOT"; do
  CAV_NVCC_EXTRA="$f" python -c "from adrates_b200 import build as b; b.build(force=True)" || exit 1
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_units_mma --csv --log-file /tmp/l.csv python tools/units_time.py $n private > /dev/null 2>&1
  echo "[classes: $f]"
  python - <<'P'
import csv
rows=[r for r in csv.reader(open('/tmp/l.csv')) if len(r)>5 and r[0].isdigit()]
import collections
d=collections.OrderedDict()
for r in rows:
    name=r[4].split('(')[0]; v=float(r[-1].replace(',',''))
    d.setdefault(name,[]).append(v)
for k,v in d.items(): print("  ",k,"min",min(v),"n",len(v), r[-2] if False else "")
print("   unit of time:", rows[0][-2] if rows else None)
P
done
CAV_NVCC_EXTRA="" python -c "from adrates_b200 import build as b; b.build(force=True)"
