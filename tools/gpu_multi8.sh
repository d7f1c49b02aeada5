#!/bin/bash
# 8-GPU: sharded parity tests + torchrun bench at 8 GPUs (N = 4M) in both exchange modes, 4 and 2 GPUs
timeout 900 python -m pytest tests/test_multigpu_gpu.py -x -q 2>&1 | tail -3
run() { g=$1; mode=$2; NBODY_EXCHANGE=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2961$g bench.py --gpus $g --steps 3 --warmup 3 > gpurun_out/bench_${g}gpu_$mode.json 2> gpurun_out/bench_${g}gpu_$mode.err; tail -1 gpurun_out/bench_${g}gpu_$mode.err | cut -c1-200; python -c "
import json,sys
d=json.load(open('gpurun_out/bench_${g}gpu_$mode.json'))
print('$g GPUs $mode:', round(d['value'],1), 'G/s', round(d['ms_per_step'],2),'ms/step', d['config']['kernel'], 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'], d['clocks'])
"; }
run 8 p2p
run 8 nccl
run 4 p2p
run 2 p2p
NBODY_GPUS=8 timeout 300 ./cuda-to-sycl-nbody_b200/bin/nbody_b200 16384 1 0.999998 0.005 1e-7 2.0 5
