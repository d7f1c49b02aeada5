#!/bin/bash
# ncu: warp-streaming vs CTA-tiled packed kernels at N = 1M (run under gpurun)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:force_wstream -s 1 -c 1 -o gpurun_out/prof_wstream_r4b32_1m \
    python tools/run_steps.py --n 1048576 --kernel auto --cfg 4,32,3 --steps 2 > gpurun_out/ncu_w1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:force_packed -s 1 -c 1 -o gpurun_out/prof_packed_r4b256_1m \
    python tools/run_steps.py --n 1048576 --kernel packed --cfg 4,256,1 --steps 2 > gpurun_out/ncu_p1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:force_wstream -s 1 -c 1 -o gpurun_out/prof_wstream_r4b128_1m \
    python tools/run_steps.py --n 1048576 --kernel auto --cfg 4,128,3 --steps 2 > gpurun_out/ncu_w2.log 2>&1
tail -2 gpurun_out/ncu_w1.log gpurun_out/ncu_p1.log gpurun_out/ncu_w2.log
