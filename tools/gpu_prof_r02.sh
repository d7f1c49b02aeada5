#!/bin/bash
# Round-2 evidence on ONE B200 (run under gpurun): ncu full captures of the production kernels at the sizes the
# bench lines are quoted on, the launch list of the bench command, compute-sanitizer on the ticketed hand-off.
# Outputs go to gpurun_out/ (scratch); tools/summarize_ncu.py turns the .ncu-rep files into profiles/r02_*.txt.
mkdir -p gpurun_out
cap() {  # name, kernel regex, args of run_steps.py
  name=$1; rx=$2; shift 2
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 1 -o gpurun_out/r02_prof_$name \
      python tools/run_steps.py "$@" > gpurun_out/r02_ncu_$name.log 2>&1
  tail -n 1 gpurun_out/r02_ncu_$name.log | cut -c1-160
}
cap wseg_r6_1m force_wseg --n 1048576 --steps 2
cap wseg_r2_262144 force_wseg --n 262144 --steps 2
cap wseg_r6_shard8_4m force_wseg --n 524288 --steps 2          # the per-GPU shard size of configs[3] at 8 GPUs (j-range differs: see DESIGN)
cap wscalar_r1_12800 force_wscalar --n 12800 --steps 2 --iters 4
cap wscalar_r1_51200 force_wscalar --n 51200 --steps 2 --iters 4
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-kernel > gpurun_out/r02_bench_under_ncu.log 2>&1
tail -n 1 gpurun_out/r02_bench_under_ncu.log | cut -c1-120
# sanitizers: the ticketed j-segment hand-off with >= 4 segments at N >= 262144 (memcheck), racecheck on a smaller problem
NBODY_SEGS=6 timeout 600 compute-sanitizer --tool memcheck python tools/run_steps.py --n 262144 --steps 1 > gpurun_out/r02_sanitizer_memcheck_seg.log 2>&1; tail -n 2 gpurun_out/r02_sanitizer_memcheck_seg.log
NBODY_SEGS=4 timeout 900 compute-sanitizer --tool racecheck python tools/run_steps.py --n 65536 --cfg 2,32,4 --steps 1 > gpurun_out/r02_sanitizer_racecheck_seg.log 2>&1; tail -n 2 gpurun_out/r02_sanitizer_racecheck_seg.log
timeout 600 compute-sanitizer --tool memcheck python tools/run_steps.py --n 5000 --steps 2 --iters 2 > gpurun_out/r02_sanitizer_memcheck.log 2>&1; tail -n 2 gpurun_out/r02_sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python tools/run_steps.py --n 5000 --steps 2 --iters 2 > gpurun_out/r02_sanitizer_racecheck.log 2>&1; tail -n 2 gpurun_out/r02_sanitizer_racecheck.log
