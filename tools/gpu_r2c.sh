#!/bin/bash
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bench_shapes.py -m gpu -q --timeout 900 -p no:cacheprovider -s -k "not cnet" 2>&1 | grep -E "worst five|passed|failed|Error" | cut -c1-900 > gpurun_out/pytest_shapes_$TAG.log
timeout 120 python tools/tc_prof.py p > gpurun_out/tcprof_p_$TAG.txt 2>&1
timeout 120 python tools/tc_prof.py ps > gpurun_out/tcprof_ps_$TAG.txt 2>&1
timeout 300 python bench.py --no-cpu-baseline --no-extra > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_wgrad_kernel" -c 2 --csv --log-file gpurun_out/wg_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/ncu_wg_$TAG.log 2>&1
cat gpurun_out/pytest_shapes_$TAG.log; cat gpurun_out/tcprof_p_$TAG.txt gpurun_out/tcprof_ps_$TAG.txt; cut -c1-330 gpurun_out/bench_$TAG.json; grep -v "^==" gpurun_out/wg_$TAG.csv | tail -3
