#!/bin/bash
# Final evidence run (1 GPU): full gpu test suite, smoke, bench (+reference arm) with a clocks log, ncu launch lists for every
# BASELINE configuration, ncu --set full captures of the three tensor-core kernels, phase map, tile timeline.
TAG=${1:-r2z}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/pytest_$TAG.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_$TAG.csv &
SMI=$!
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
kill $SMI

timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/ncu_list_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"tc_cell_fwd_kernel" -s 30 -c 4 -o gpurun_out/prof_tc_$TAG -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/ncu_tc_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"tc_cell_bwd_kernel|tc_wgrad_kernel" -s 58 -c 3 -o gpurun_out/prof_wg_$TAG -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/ncu_wg_$TAG.log 2>&1
timeout 120 python tools/phase_times.py > gpurun_out/phases_$TAG.txt 2>&1
timeout 120 python tools/tc_prof.py ps > gpurun_out/tcprof_ps_$TAG.txt 2>&1
timeout 120 python tools/tc_prof.py p > gpurun_out/tcprof_p_$TAG.txt 2>&1
cat gpurun_out/pytest_$TAG.log; tail -2 gpurun_out/smoke_$TAG.log; cut -c1-400 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err; cat gpurun_out/phases_$TAG.txt
