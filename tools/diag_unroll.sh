#!/bin/bash
# usage: tools/diag_unroll.sh   units stage of the dedup 1M-trade book (latency-bound: 2 900 tiles) for K-loop unroll depths
for f in "" "-DMMA_UNROLL=4" "-DMMA_UNROLL=8" "-DMMA_UNROLL=1"; do
  CAV_NVCC_EXTRA="$f" python -c "from adrates_b200 import build as b; b.build(force=True)" || exit 1
  echo "[build: $f]"
  python tools/units_time.py 1000000 dedup 2>&1 | grep "units-stage"
  CAV_UNITS_WS=0 python tools/units_time.py 1000000 dedup 2>&1 | grep "units-stage" | sed 's/^/  WS=0 /'
  CAV_UNITS_WS=1 python tools/units_time.py 1000000 dedup 2>&1 | grep "units-stage" | sed 's/^/  WS=1 /'
done
CAV_NVCC_EXTRA="" python -c "from adrates_b200 import build as b; b.build(force=True)"
