#!/bin/bash
# usage: tools/diag_bsmem.sh [n]   upper bound of staging the B operands of the tile kernel's K loop in shared memory:
# units stage of the private layout and per-class durations for the normal build and -DMMA_DIAG_BSMEM (B read from shared
# memory with a conflict-free pattern: wrong values, timing only)
n=${1:-300000}
for f in "" "-DMMA_DIAG_BSMEM" "-DMMA_DIAG_BSMEM -DMMA_DIAG_NOSTORE" "-DMMA_DIAG_BHOT"; do
  CAV_NVCC_EXTRA="$f" python -c "from adrates_b200 import build as b; b.build(force=True)" || exit 1
  echo "[build: $f]"
  python tools/units_time.py $n private 2>&1 | grep "units-stage"
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_units_mma --csv --log-file /tmp/l.csv python tools/units_time.py $n private > /dev/null 2>&1
  python - <<'P'
import csv, collections
rows=[r for r in csv.reader(open('/tmp/l.csv')) if len(r)>5 and r[0].isdigit()]
d=collections.OrderedDict()
for r in rows:
    d.setdefault(r[4].split('(')[0],[]).append(float(r[-1].replace(',','')))
for k,v in d.items(): print("  ",k,"min",min(v)/1000,"us")
P
done
CAV_NVCC_EXTRA="" python -c "from adrates_b200 import build as b; b.build(force=True)"
