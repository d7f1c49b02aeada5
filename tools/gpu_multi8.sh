#!/bin/bash
# 8-GPU: sharded parity tests + torchrun bench at 8 and 4 GPUs (N = 4M)
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests/test_multigpu_gpu.py -x -q 2>&1 | tail -4
for g in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2951$g bench.py --gpus $g --steps 3 --warmup 3 > gpurun_out/bench_${g}gpu.json 2> gpurun_out/bench_${g}gpu.err
tail -2 gpurun_out/bench_${g}gpu.err; cut -c1-420 gpurun_out/bench_${g}gpu.json
done
NBODY_GPUS=8 timeout 300 ./cuda-to-sycl-nbody_b200/bin/nbody_b200 16384 1 0.999998 0.005 1e-7 2.0 5
