#!/bin/bash
# Round profile set (run under gpurun, 1 GPU).  Outputs land in gpurun_out/ and are summarised into profiles/ afterwards.
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 600 gpurun_out/bench_n1.json
# every launch of two steps after one warm-up step (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 3 > gpurun_out/ncu_launch.log 2>&1
# full captures of the heavy kernels (one launch each, second step)
for spec in "rows_gemm_ws:12:1:full_rows_gemm_fwd" "rows_gemm_ws:17:1:full_rows_gemm_dgrad" "wgrad_tc:7:1:full_wgrad" "colstat4:7:1:full_colstat" "norm_bwd_apply4:6:1:full_nba" "pairwise_bce:1:1:full_loss"; do
  IFS=: read k s c o <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:"$k" -s "$s" -c "$c" -o gpurun_out/"$o" -f python tools/profile_step.py 2 > gpurun_out/"$o".log 2>&1
  ncu -i gpurun_out/"$o".ncu-rep --page raw --csv > gpurun_out/"$o"_raw.csv 2>/dev/null
done
ls -la gpurun_out | head -30
