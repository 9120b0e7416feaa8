for f in "-DGW_NSTORE=0" "-DGW_NSTORE=0 -DGW_NBUF=3" "-DGW_NSTORE=0 -DMMA_UNROLL=4" "-DGW_NSTORE=0 -DMMA_UNROLL=1" "-DGW_NSTORE=6 -DGW_NBUF=2"; do
  CAV_NVCC_EXTRA="$f" python -c "from adrates_b200 import build as b; b.build(force=True)" || exit 1
  echo "[$f]"; CAV_UNITS_WS=1 timeout 200 python tools/units_time.py 300000 private 2>&1 | grep " units "
done
