#!/bin/bash
# final-kernel evidence: ncu full capture of the production kernel at N=1M, launch list of the bench
# command, compute-sanitizer memcheck + racecheck on a small problem
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:force_wseg -s 1 -c 1 -o gpurun_out/prof_wseg_r6_1m \
    python tools/run_steps.py --n 1048576 --kernel auto --steps 2 > gpurun_out/ncu_wseg.log 2>&1
tail -n 2 gpurun_out/ncu_wseg.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-kernel > gpurun_out/bench_under_ncu.log 2>&1
tail -n 1 gpurun_out/bench_under_ncu.log | cut -c1-160
compute-sanitizer --tool memcheck python tools/run_steps.py --n 5000 --steps 2 --iters 2 > gpurun_out/sanitizer_memcheck.log 2>&1; tail -n 3 gpurun_out/sanitizer_memcheck.log
compute-sanitizer --tool racecheck python tools/run_steps.py --n 5000 --steps 2 --iters 2 > gpurun_out/sanitizer_racecheck.log 2>&1; tail -n 3 gpurun_out/sanitizer_racecheck.log
NBODY_SEGS=4 compute-sanitizer --tool memcheck python tools/run_steps.py --n 70000 --steps 1 --iters 2 > gpurun_out/sanitizer_memcheck_seg.log 2>&1; tail -n 3 gpurun_out/sanitizer_memcheck_seg.log
