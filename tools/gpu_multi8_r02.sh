#!/bin/bash
# Round-2 8-GPU evidence (gpurun --gpus 8; kept short: an 8-GPU lease costs 8x GPU-minutes): the P = 8 parity tests after the
# last kernel change and one 8-GPU bench line (N = 4194304, torchrun, parity key).
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -rs -k "8" > gpurun_out/r02_pytest_multigpu8.txt 2>&1; tail -n 6 gpurun_out/r02_pytest_multigpu8.txt
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 4 --warmup 3 --no-same-n \
    > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_8gpu.json')); print(8, d['value'], d['pct_fp32_roofline'], d['e2e']['value'], d['parity']['matches_reference_golden'], d['config']['kernel'])"
