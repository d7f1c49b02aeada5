#!/bin/bash
# Round-2 call 4 on ONE B200 (under gpurun): GPU test-suite, AUTO switch-point sweep of the rebuilt library, bench line.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.txt 2>&1; tail -n 3 gpurun_out/r02_pytest_gpu.txt
{
  for n in 32768 40000 50000 57720 65536 80000 100000 115000 131072 160000 200000 227328 262144 300000 340992 400003 524288; do
    SWEEP_FAMILY=6 python tools/sweep_cfg.py $n 1 0 5
    python tools/sweep_cfg.py $n 2,4,6 0 5
  done
} > gpurun_out/r02_sched_sweep_raw.txt 2>&1
tail -n 8 gpurun_out/r02_sched_sweep_raw.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_1gpu.json')); print(d['value'], d['pct_fp32_roofline'], d['e2e']['value'], d['parity']['matches_reference_golden'], d['config']['kernel'], d['clocks'])"
