#!/bin/bash
# quick check: backward parity (small + bench shape NC), wgrad/bwd/fwd kernel times, bench
TAG=${1:-q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_vec.py tests/test_gpu_policy.py -m gpu -q --timeout 600 -p no:cacheprovider -x 2>&1 | tail -4 > gpurun_out/pytest_$TAG.log
timeout 900 python -m pytest tests/test_gpu_bench_shapes.py -m gpu -q --timeout 900 -p no:cacheprovider -s -k "nc_catchup or cnet" 2>&1 | grep -E "worst five|passed|failed|Error" | cut -c1-600 > gpurun_out/pytest_shapes_$TAG.log
timeout 300 python bench.py --no-cpu-baseline --no-extra > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/ncu_list_$TAG.log 2>&1
timeout 120 python tools/phase_times.py > gpurun_out/phases_$TAG.txt 2>&1
cat gpurun_out/pytest_$TAG.log gpurun_out/pytest_shapes_$TAG.log; cut -c1-330 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err; cat gpurun_out/phases_$TAG.txt
python tools/summarize_ncu.py $TAG 2>/dev/null | head -9
