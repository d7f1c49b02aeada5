#!/bin/bash
# Round-2 evidence on ONE B200 (run under gpurun): GPU test-suite, ncu full captures of the production kernels at the sizes
# the bench lines are quoted on, the launch list of the bench command, compute-sanitizer on the generated tile bodies and
# the ticketed hand-off, the small-N table against the reference kernel, the final bench line.
# Outputs go to gpurun_out/ (scratch); tools/summarize_ncu.py turns the .ncu-rep files into profiles/r02_*.txt.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.txt 2>&1; tail -n 2 gpurun_out/r02_pytest_gpu.txt
cap() {  # name, kernel regex, args of run_steps.py
  name=$1; rx=$2; shift 2
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 1 -o gpurun_out/r02_prof_$name \
      python tools/run_steps.py "$@" > gpurun_out/r02_ncu_$name.log 2>&1
  tail -n 1 gpurun_out/r02_ncu_$name.log | cut -c1-160
}
cap wseg_r6_1m_gen force_wseg --n 1048576 --steps 2
cap wseg_r4_262144_gen force_wseg --n 262144 --steps 2
cap wseg_r2_131072_gen force_wseg --n 131072 --steps 2
cap wseg_r6_shard8_4m_gen force_wseg --n 524288 --steps 2          # the per-GPU shard size of configs[3] at 8 GPUs (j-range differs: see DESIGN)
cap wscalar_r1_12800 force_wscalar --n 12800 --steps 2 --iters 4
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-kernel > gpurun_out/r02_bench_under_ncu.log 2>&1
tail -n 1 gpurun_out/r02_bench_under_ncu.log | cut -c1-120
python tools/small_n.py > gpurun_out/r02_small_n.txt 2>&1; cat gpurun_out/r02_small_n.txt
# sanitizers: generated tile bodies + ticketed j-segment hand-off with >= 4 segments at N >= 262144 (memcheck), racecheck smaller
NBODY_SEGS=6 timeout 600 compute-sanitizer --tool memcheck python tools/run_steps.py --n 262144 --steps 1 > gpurun_out/r02_sanitizer_memcheck_seg.log 2>&1; tail -n 2 gpurun_out/r02_sanitizer_memcheck_seg.log
NBODY_SEGS=4 timeout 900 compute-sanitizer --tool racecheck python tools/run_steps.py --n 65536 --cfg 2,32,4 --steps 1 > gpurun_out/r02_sanitizer_racecheck_seg.log 2>&1; tail -n 2 gpurun_out/r02_sanitizer_racecheck_seg.log
timeout 600 compute-sanitizer --tool memcheck python tools/run_steps.py --n 5000 --steps 2 --iters 2 > gpurun_out/r02_sanitizer_memcheck.log 2>&1; tail -n 2 gpurun_out/r02_sanitizer_memcheck.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
python bench.py --steps 10 --warmup 3 --bodies 262144 --no-cpu-baseline > gpurun_out/r02_bench_1gpu_262144.json 2> /dev/null
python -c "
import json
for f in ('r02_bench_1gpu','r02_bench_1gpu_262144'):
    d=json.load(open('gpurun_out/'+f+'.json')); print(f, d['value'], d['pct_fp32_roofline'], d['e2e']['value'], d['parity']['matches_reference_golden'], d['config']['kernel'], d.get('reference_cuda_kernel',{}).get('value'))"
