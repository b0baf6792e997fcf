#!/bin/bash
# GPU box: what the driver runs at round end (tests, smoke, both bench arms), plus the ncu launch list.
TAG=${1:-r01e}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu --durations=6 2>&1 | tail -12 > gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.log 2>&1; tail -1 gpurun_out/${TAG}_bench_reference.log | cut -c1-300
python bench.py > gpurun_out/${TAG}_bench.log 2>&1; tail -1 gpurun_out/${TAG}_bench.log | cut -c1-2500
# same workload and batch as the bench line (the mapping of the token parse depends on the stream count); fewer distinct
# images and steps only shorten the corpus build and the serialised profiling run
CMD="python bench.py --distinct 32 --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
tail -2 gpurun_out/${TAG}_ncu1.log
