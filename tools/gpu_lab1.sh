#!/bin/bash
# Round-2 lab run on ONE B200 (under gpurun): the restored bench line, then the SASS-level experiments of
# tools/sass_lab.py (synthetic tile bodies at 4/8/12/16 resident warps per SM) and the real schedules.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/lab1_gpu.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/lab1_bench.json 2> gpurun_out/lab1_bench.err
tail -c 600 gpurun_out/lab1_bench.json
{
  echo "## real schedules (parity checked against the reference golden at N=1M)"
  for f in lab_build/real_*.so; do
    printf "%-40s " $(basename $f); NBODY_LAB_PARITY=1 timeout 120 python tools/lab_one.py $f 1048576 3 || echo FAILED
  done
  echo "## synthetic blocks"
  python tools/sass_lab.py build lab_build/base_w16.so lab_build/exp
  python tools/sass_lab.py build lab_build/base_w8.so lab_build/exp --subset
  python tools/sass_lab.py build lab_build/base_w4.so lab_build/exp --subset
  python tools/sass_lab.py run lab_build/exp --bodies 1048576
} > gpurun_out/lab1.txt 2>&1
tail -n 70 gpurun_out/lab1.txt
