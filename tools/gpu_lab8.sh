#!/bin/bash
# Round-2 call 8 on ONE B200: the accumulator-relay kernel (family 7) -- parity against the scalar kernel at every shape
# and size, per-body masses, multi-iteration states, times; memcheck / racecheck of one small launch.
mkdir -p gpurun_out
LAB_BUDGET_S=90 timeout 150 python tools/lab_relay.py 2>&1 | cut -c1-260
export NBODY_KERNEL_CONFIG=16,128,7
timeout 60 compute-sanitizer --tool memcheck python tools/lab_one.py cuda-to-sycl-nbody_b200/lib/libnbody_b200.so 2049 1 > gpurun_out/lab8_memcheck.txt 2>&1; tail -n 2 gpurun_out/lab8_memcheck.txt
timeout 60 compute-sanitizer --tool racecheck python tools/lab_one.py cuda-to-sycl-nbody_b200/lib/libnbody_b200.so 1000 1 > gpurun_out/lab8_racecheck.txt 2>&1; tail -n 2 gpurun_out/lab8_racecheck.txt
