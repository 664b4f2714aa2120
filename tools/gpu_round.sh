#!/bin/bash
# One GPU session: smoke, bench, ncu launch list, ncu full capture of the dominant kernel.
# Usage (under gpurun):  bash tools/gpu_round.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke_$TAG.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_$TAG.csv &
SMI=$!
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
kill $SMI
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cell_fwd_kernel -s 20 -c 2 -o gpurun_out/prof_fwd_$TAG -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_fwd_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cell_bwd_kernel|wgrad_kernel" -s 4 -c 2 -o gpurun_out/prof_bwd_$TAG -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bwd_$TAG.log 2>&1
tail -3 gpurun_out/smoke_$TAG.log; cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
