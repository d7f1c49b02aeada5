#!/bin/bash
# ncu captures of the two main kernel families + clock sampling under load (run under gpurun)
set -x
mkdir -p gpurun_out
Q=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap
nvidia-smi --query-gpu=$Q --format=csv -lms 200 > gpurun_out/clocks_run.csv &
SMI=$!
python tools/run_steps.py --n 1048576 --kernel packed --cfg 4,256 --steps 8
python tools/run_steps.py --n 1048576 --kernel scalar --cfg 4,256 --steps 4
kill $SMI
ncu --set full --clock-control none --import-source on -k regex:force_packed -s 1 -c 1 -o gpurun_out/prof_packed_r4b256 \
    python tools/run_steps.py --n 262144 --kernel packed --cfg 4,256 --steps 2 > gpurun_out/ncu_packed.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:force_scalar -s 1 -c 1 -o gpurun_out/prof_scalar_r4b256 \
    python tools/run_steps.py --n 262144 --kernel scalar --cfg 4,256 --steps 2 > gpurun_out/ncu_scalar.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:force_packed -s 1 -c 1 -o gpurun_out/prof_packed_r2b128 \
    python tools/run_steps.py --n 262144 --kernel packed --cfg 2,128 --steps 2 > gpurun_out/ncu_packed2.log 2>&1
tail -3 gpurun_out/ncu_packed.log
ls -la gpurun_out
