set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t_full.log 2>&1; tail -4 gpurun_out/t_full.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 1500 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 600 gpurun_out/bench_reference.json
