#!/bin/bash
TAG=${1:-r2d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_policy.py tests/test_gpu_backward.py tests/test_gpu_vec.py tests/test_gpu_hetero.py -m gpu -q --timeout 600 -p no:cacheprovider -x 2>&1 | tail -8 > gpurun_out/pytest_$TAG.log
timeout 900 python -m pytest tests/test_gpu_bench_shapes.py -m gpu -q --timeout 900 -p no:cacheprovider -s 2>&1 | grep -E "worst five|passed|failed|Error" | cut -c1-900 > gpurun_out/pytest_shapes_$TAG.log
timeout 120 python tools/tc_prof.py p > gpurun_out/tcprof_p_$TAG.txt 2>&1
timeout 120 python tools/tc_prof.py ps > gpurun_out/tcprof_ps_$TAG.txt 2>&1
timeout 300 python bench.py --no-cpu-baseline --no-extra > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/ncu_list_$TAG.log 2>&1
timeout 120 python tools/phase_times.py > gpurun_out/phases_$TAG.txt 2>&1
cat gpurun_out/pytest_$TAG.log gpurun_out/pytest_shapes_$TAG.log; cat gpurun_out/tcprof_p_$TAG.txt gpurun_out/tcprof_ps_$TAG.txt | grep -v "^  q=.[^0-9]"; cut -c1-330 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err; cat gpurun_out/phases_$TAG.txt
