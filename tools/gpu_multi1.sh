#!/bin/bash
# 2-GPU bring-up: sharded parity tests, torchrun bench at N=2, single-process scaling probe
timeout 600 python -m pytest tests/test_multigpu_gpu.py -x -q 2>&1 | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -5 gpurun_out/bench_2gpu.err; cat gpurun_out/bench_2gpu.json
timeout 300 python tools/run_steps.py --n 4194304 --gpus 2 --steps 2 --iters 2
