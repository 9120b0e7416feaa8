#!/bin/bash
# usage: tools/variant_time.sh "<nvcc -D flags>" <tag> [n] [layout]   (rebuild on the GPU box, time the units stage)
flags="$1"; tag="$2"; shift 2
CAV_NVCC_EXTRA="$flags" python -c "from adrates_b200 import build as b; b.build(force=True)" || exit 1
python tools/units_time.py "$@" 2>&1 | grep " units " | tail -1 | sed "s/^/[$tag: $flags] /"
