#!/bin/bash
# (1) ncu launch list of the bench command (per-launch device time; shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-kernel > gpurun_out/bench_under_ncu.log 2>&1
tail -1 gpurun_out/bench_under_ncu.log | cut -c1-200
# (2) shard-sized problem (N/P = 524288 bodies): residency and R
for k in 0 24 28 32; do python tools/run_steps.py --n 524288 --kernel auto --cfg 4,32,3 --resident $k --steps 3 | tail -1 | cut -c1-200 | sed "s/^/k=$k /"; done
for k in 0 28 32; do python tools/run_steps.py --n 524288 --kernel auto --cfg 2,32,3 --resident $k --steps 3 | tail -1 | cut -c1-200 | sed "s/^/k=$k /"; done
python tools/run_steps.py --n 524288 --kernel auto --cfg 4,64,3 --steps 3 | tail -1 | cut -c1-200
python tools/run_steps.py --n 524288 --kernel packed --cfg 2,128,1 --steps 3 | tail -1 | cut -c1-200
python tools/run_steps.py --n 524288 --kernel packed --cfg 4,128,1 --steps 3 | tail -1 | cut -c1-200
python tools/run_steps.py --n 4194304 --kernel auto --steps 2 | tail -1 | cut -c1-200
