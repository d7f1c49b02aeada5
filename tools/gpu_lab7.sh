#!/bin/bash
# Round-2 call 7 on ONE B200: ncu source-level captures of the scalar small-shard kernel at N = 12800 (one warp per
# sub-partition), ptxas' code (V0) and the best generated variant (Sm): where does the single warp wait?
mkdir -p gpurun_out
export NBODY_KERNEL_CONFIG=1,32,6
for v in V0 Sm; do
  timeout 100 ncu --set full --clock-control none --import-source on -k regex:force_wscalar -s 1 -c 1 -f -o gpurun_out/r02_prof_wscalar_$v \
     python tools/lab_one.py lab_build/s/$v.so 12800 2 > gpurun_out/r02_ncu_wscalar_$v.log 2>&1
  tail -n 1 gpurun_out/r02_ncu_wscalar_$v.log | cut -c1-160
done
ls -la gpurun_out/*.ncu-rep
