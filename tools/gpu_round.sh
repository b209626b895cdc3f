#!/bin/bash
# One gpurun call for a whole validation pass (box acquisition + snapshot push cost 1-3 GPU-minutes per call, so batch):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_round.sh [tests] [bench] [ops] [launches] [full]'
# Each stage has its own timeout so a hung kernel cannot hold the box; outputs land in gpurun_out/.
mkdir -p gpurun_out
stages="${*:-tests bench}"
for s in $stages; do
  case "$s" in
    tests)    timeout 900 python -m pytest tests -q -m gpu --timeout 200 --tb=short -x --maxfail=25 > gpurun_out/pytest_gpu.log 2>&1; grep -v "Warning" gpurun_out/pytest_gpu.log | tail -60 ;;
    testsall) timeout 1200 python -m pytest tests -q -m gpu --timeout 200 --tb=short --maxfail=40 > gpurun_out/pytest_gpu.log 2>&1; grep -v "Warning" gpurun_out/pytest_gpu.log | tail -150
              if ! tail -1 gpurun_out/pytest_gpu.log | grep -q " passed" || tail -1 gpurun_out/pytest_gpu.log | grep -q failed; then
                echo "== bisect: scorer tests with the regular last-chunk mapping (PTRB200_NO_PARTIAL=1)"
                PTRB200_NO_PARTIAL=1 PTRB200_NO_RUNS=1 timeout 600 python -m pytest tests/test_gpu_scorer.py tests/test_gpu_tc_gemm.py tests/test_gpu_r2_parity.py tests/test_gpu_losses.py -q -m gpu --timeout 200 --tb=line 2>&1 | grep -v Warning | tail -15
              fi ;;
    smoke)    timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ;;
    bench)    timeout 200 python bench.py 2>gpurun_out/bench_n1.err | tail -1 > gpurun_out/bench_n1.json
              python -c "import json; d=json.loads(open('gpurun_out/bench_n1.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']); print(d['roofline']['kernels_ms_per_step'])" ;;
    ops)      timeout 300 python tools/op_bench.py > gpurun_out/op_bench.log 2>&1; cp profiles/r0*_op_table.md gpurun_out/; tail -42 gpurun_out/r0*_op_table.md ;;
    launches) timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 3 > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log ;;
    full)     timeout 800 bash tools/profile_all.sh 2>&1 | tail -3 ;;
    *)        echo "unknown stage $s" ;;
  esac
done
