#!/bin/bash
# usage: tools/diag_units.sh [n]   phase breakdown of k_units_mma on the private layout by diagnostic builds (run on the GPU box)
n=${1:-500000}
for f in "" "-DMMA_DIAG_NOSTORE" "-DMMA_DIAG_NOMMA" "-DMMA_DIAG_BHOT" "-DMMA_DIAG_NOSCAL" "-DMMA_DIAG_NOSTORE -DMMA_DIAG_NOMMA" "-DMMA_DIAG_NOSTORE -DMMA_DIAG_NOMMA -DMMA_DIAG_NOSCAL" "-DMMA_DIAG_NOSTORE -DMMA_DIAG_BHOT -DMMA_DIAG_NOSCAL"; do
  bash tools/variant_time.sh "$f" diag $n private
done
CAV_NVCC_EXTRA="" python -c "from adrates_b200 import build as b; b.build(force=True)"
