#!/bin/bash
# usage: tools/variant_bench.sh "<nvcc -D flags>" <tag> [bench args]   (rebuilds the library on the GPU box, runs the private-layout bench)
flags="$1"; tag="$2"; shift 2
CAV_NVCC_EXTRA="$flags" python -c "from adrates_b200 import build as b; b.build(force=True)" || exit 1
python bench.py --steps 5 --warmup 3 --no-extra --layout private --cpu-sample 100 "$@" > gpurun_out/v_$tag.json 2> gpurun_out/v_$tag.err
tail -2 gpurun_out/v_$tag.err | cut -c1-200; python tools/show_bench.py gpurun_out/v_$tag.json | sed "s/^/[$tag: $flags] /"
