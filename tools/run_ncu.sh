#!/bin/bash
# usage: tools/run_ncu.sh <kernel-regex> <skip> <count> <outname>   (run under gpurun; never a bench value)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$1" -s "$2" -c "$3" -o gpurun_out/"$4" -f python tools/profile_step.py 2 > gpurun_out/"$4".log 2>&1
ncu -i gpurun_out/"$4".ncu-rep --page source --csv > gpurun_out/"$4"_src.csv 2>/dev/null
tail -2 gpurun_out/"$4".log
