#!/bin/bash
# Round-2 GPU session B (1 GPU): new kernels (4-slot A ring, PDL, bias partials, h_seq redirect, env v2) -- parity first,
# then A/B benches (PDL on/off, RAW tiles on/off), ncu launch list + full captures, phase map.
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tc.py tests/test_gpu_env.py tests/test_gpu_policy.py tests/test_gpu_backward.py tests/test_gpu_vec.py tests/test_gpu_hetero.py \
    -m gpu -q --timeout 600 -p no:cacheprovider -x 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log
timeout 1500 python -m pytest tests/test_gpu_bench_shapes.py -m gpu -q --timeout 900 -p no:cacheprovider -s 2>&1 | tail -30 > gpurun_out/pytest_shapes_$TAG.log
NMARL_RAW_TILES=1 timeout 1500 python -m pytest tests/test_gpu_backward.py tests/test_gpu_bench_shapes.py -m gpu -q --timeout 900 -p no:cacheprovider -s 2>&1 | tail -30 > gpurun_out/pytest_raw_$TAG.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1
for v in base nopdl raw; do
  case $v in base) E="";; nopdl) E="NMARL_NO_PDL=1";; raw) E="NMARL_RAW_TILES=1";; esac
  env $E timeout 300 python bench.py --no-cpu-baseline --no-extra > gpurun_out/bench_${v}_$TAG.json 2> gpurun_out/bench_${v}_$TAG.err
done
NMARL_RAW_TILES=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/ncu_list_$TAG.log 2>&1
NMARL_RAW_TILES=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"tc_cell_fwd_kernel" -s 30 -c 4 -o gpurun_out/prof_tc_$TAG -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/ncu_tc_$TAG.log 2>&1
NMARL_RAW_TILES=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"tc_cell_bwd_kernel|tc_wgrad_kernel|cacc_step" -s 58 -c 4 -o gpurun_out/prof_wg_$TAG -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/ncu_wg_$TAG.log 2>&1
NMARL_RAW_TILES=1 timeout 120 python tools/phase_times.py > gpurun_out/phases_$TAG.txt 2>&1
for f in pytest pytest_shapes pytest_raw smoke phases; do echo "== $f"; tail -12 gpurun_out/${f}_$TAG.*; done
for v in base nopdl raw; do echo "== bench $v"; cut -c1-330 gpurun_out/bench_${v}_$TAG.json; tail -2 gpurun_out/bench_${v}_$TAG.err; done
