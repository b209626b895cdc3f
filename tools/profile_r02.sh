#!/bin/bash
# Round-2 profile set (run under gpurun, 1 GPU; nothing printed by a run under ncu is a bench value).
#   bash tools/profile_r02.sh [launches] [b] [c] [d] [e]
# Outputs land in gpurun_out/ (full_<cfg>_<name>.ncu-rep + _raw.csv) and are summarised into profiles/ by
# tools/summarise_profiles.py r02.
mkdir -p gpurun_out
cap() {   # cap <cfg> <kernel-regex> <skip> <name>: one launch of the kernel from the second step
  local cfg=$1 k=$2 s=$3 o=full_$1_$4
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$k" -s "$s" -c 1 -o gpurun_out/"$o" -f python tools/profile_step.py 2 0 "$cfg" > gpurun_out/"$o".log 2>&1
  ncu -i gpurun_out/"$o".ncu-rep --page raw --csv > gpurun_out/"$o"_raw.csv 2>/dev/null
  ncu -i gpurun_out/"$o".ncu-rep --page source --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin))
w = csv.writer(sys.stdout)
for r in rows: w.writerow(r[:6])" > gpurun_out/"$o"_source.csv
  rm -f gpurun_out/"$o".ncu-rep          # the reports (with imported sources) exceed what a gpurun call brings back; the two CSV pages are what gets read
  tail -1 gpurun_out/"$o".log
}
for stage in "${@:-launches b c d e}"; do
  case "$stage" in
    launches)
      timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 3 > gpurun_out/ncu_launch.log 2>&1
      ENC_LAYERS=3 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_c.csv python tools/profile_step.py 2 0 c > gpurun_out/ncu_launch_c.log 2>&1
      tail -1 gpurun_out/ncu_launch.log gpurun_out/ncu_launch_c.log ;;
    b)  # headline: one hidden forward layer, one dgrad, one wgrad, the dY sweep, the loss (launch indices of step 2)
      cap b rows_gemm_ws_kernel 12 rows_gemm_fwd
      cap b rows_gemm_ws_kernel 17 rows_gemm_dgrad
      cap b wgrad_tc 7 wgrad
      cap b colstat4 6 colstat
      cap b "lambdarank_runs|pairwise_bce" 1 loss ;;
    c)  # list scorer (L=3): attention GEMMs (QK^T and PV), row softmax, the wide-layer kernels, ApproxNDCG
      export ENC_LAYERS=3
      cap c "bgemm_(fast|nt_tc)_kernel" 18 attn_qk
      cap c "bgemm_(fast|nt_tc)_kernel" 19 attn_pv
      cap c softmax_rows_kernel 3 softmax
      cap c softmax_bwd_rows_kernel 3 softmax_bwd
      cap c rows_gemm_tc_kernel 24 rows_gemm_tc          # 136 -> 136 output projection of the first encoder layer
      cap c rows_gemm_tc_kernel 31 rows_gemm_tc_wide     # 256 -> 512 layer of the head net
      cap c wgrad_tc 15 wgrad                            # its weight gradient (column-blocked)
      cap c approxndcg_kernel 1 loss
      unset ENC_LAYERS ;;
    d)  cap d lambdaloss_kernel 1 loss ;;
    e)  cap e listmle_kernel 1 loss ;;
    *)  echo "unknown stage $stage" ;;
  esac
done
ls gpurun_out | grep -c raw.csv
