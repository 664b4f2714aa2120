#!/bin/bash
# Round-2 GPU session A (1 GPU): full gpu test suite (incl. bench-shape parity + tfnet), RAW-tile validation,
# hardware probes, smoke, bench A/B, launch list, phase map.
TAG=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu_$TAG.txt; nproc >> gpurun_out/gpu_$TAG.txt; free -g | head -2 >> gpurun_out/gpu_$TAG.txt
timeout 300 python tools/probe_tf32_operand.py > gpurun_out/probe_$TAG.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x --deselect tests/test_gpu_bench_shapes.py 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log
timeout 1500 python -m pytest tests/test_gpu_bench_shapes.py -m gpu -q --timeout 900 -p no:cacheprovider -s 2>&1 | tail -40 > gpurun_out/pytest_shapes_$TAG.log
NMARL_RAW_TILES=1 timeout 1500 python -m pytest tests/test_gpu_backward.py tests/test_gpu_vec.py tests/test_gpu_bench_shapes.py -m gpu -q --timeout 900 -p no:cacheprovider -s 2>&1 | tail -40 > gpurun_out/pytest_raw_$TAG.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
NMARL_RAW_TILES=1 timeout 300 python bench.py --no-cpu-baseline --no-extra > gpurun_out/bench_raw_$TAG.json 2> gpurun_out/bench_raw_$TAG.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/ncu_list_$TAG.log 2>&1
timeout 120 python tools/phase_times.py > gpurun_out/phases_$TAG.txt 2>&1
NMARL_RAW_TILES=1 timeout 120 python tools/phase_times.py > gpurun_out/phases_raw_$TAG.txt 2>&1
for f in probe pytest pytest_shapes pytest_raw smoke phases phases_raw; do echo "== $f"; tail -12 gpurun_out/${f}_$TAG.*; done
cut -c1-1500 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err; cut -c1-400 gpurun_out/bench_raw_$TAG.json
