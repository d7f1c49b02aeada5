#!/bin/bash
# Round-2 call 9 on ONE B200: ncu source-level captures of the accumulator-relay kernel (W = 4, TJ = 32)
mkdir -p gpurun_out
export NBODY_KERNEL_CONFIG=32,128,7
for n in 2048 12800; do
  timeout 100 ncu --set full --clock-control none --import-source on -k regex:force_wrelay -s 1 -c 1 -f -o gpurun_out/r02_prof_wrelay_4x32_$n \
     python tools/lab_one.py cuda-to-sycl-nbody_b200/lib/libnbody_b200.so $n 2 > gpurun_out/r02_ncu_wrelay_4x32_$n.log 2>&1
  tail -n 1 gpurun_out/r02_ncu_wrelay_4x32_$n.log | cut -c1-160
done
