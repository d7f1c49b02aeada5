#!/bin/bash
# Round-2 multi-GPU evidence (gpurun --gpus 4): the multi-GPU parity tests after the last kernel change, and the
# 2- and 4-GPU bench lines (torchrun, one rank per GPU) with their parity key and same-N single-GPU point.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r02_multi_gpus.txt
python -m pytest tests/test_multigpu_gpu.py -m gpu -q -rs > gpurun_out/r02_pytest_multigpu.txt 2>&1; tail -n 12 gpurun_out/r02_pytest_multigpu.txt
for g in 2 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2950$g bench.py --gpus $g --steps 5 --warmup 3 \
      > gpurun_out/r02_bench_${g}gpu.json 2> gpurun_out/r02_bench_${g}gpu.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_${g}gpu.json')); print($g, d['value'], d['pct_fp32_roofline'], d['e2e']['value'], d['parity']['matches_reference_golden'], d['config']['kernel'], d.get('same_n_scaling'))"
done
NBODY_EXCHANGE=nccl python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 3 --warmup 3 --no-same-n \
    > gpurun_out/r02_bench_4gpu_nccl.json 2> gpurun_out/r02_bench_4gpu_nccl.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_4gpu_nccl.json')); print('nccl', d['value'], d['parity']['matches_reference_golden'], d['config']['kernel'])"
