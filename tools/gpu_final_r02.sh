#!/bin/bash
# Round-2 final validation on ONE B200 (under gpurun): GPU test-suite, smoke(), per-body-mass kernel rates, racecheck of the
# hand-off with the time-out disabled (a sanitizer slows the kernel ~1000x), the final bench lines.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.txt 2>&1; tail -n 2 gpurun_out/r02_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1; tail -n 4 gpurun_out/r02_smoke.txt
{
  for n in 131072 262144 1048576; do python tools/run_steps.py --n $n --steps 3 --mass | tail -n 1; python tools/run_steps.py --n $n --steps 3 | tail -n 1; done
} > gpurun_out/r02_mass_rates.txt 2>&1; cat gpurun_out/r02_mass_rates.txt | cut -c1-200
NBODY_HANDOFF_TIMEOUT_S=0 NBODY_SEGS=4 timeout 900 compute-sanitizer --tool racecheck python tools/run_steps.py --n 16384 --cfg 2,32,4 --steps 1 > gpurun_out/r02_sanitizer_racecheck_seg.log 2>&1; tail -n 2 gpurun_out/r02_sanitizer_racecheck_seg.log
NBODY_SEGS=5 timeout 600 compute-sanitizer --tool memcheck python tools/run_steps.py --n 262144 --steps 1 --mass > gpurun_out/r02_sanitizer_memcheck_mass.log 2>&1; tail -n 2 gpurun_out/r02_sanitizer_memcheck_mass.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> /dev/null
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_1gpu.json')); print(d['value'], d['pct_fp32_roofline'], d['e2e']['value'], d['parity']['matches_reference_golden'], d['config']['kernel'], d['roofline']['issue_bound'])
r=json.load(open('gpurun_out/r02_bench_reference_arm.json')); print('reference arm', r['value'], r['cpu_baseline']['cores'])"
