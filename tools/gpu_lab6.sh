#!/bin/bash
# Round-2 sweep 6 on ONE B200 (under gpurun, ~7 GPU-minutes left): generated tile body of the scalar small-shard kernel --
# period templates / LDS distance variants against ptxas' code (tools/lab_scalar.py), then the whole GPU test-suite and
# smoke() with the best bit-exact variant in place of the production library (on the box only).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/lab6_gpu.txt 2>&1
LAB_BUDGET_S=100 timeout 150 python tools/lab_scalar.py lab_build/s/V0.so lab_build/s/S?.so 2>&1 | cut -c1-230
best=$(cat gpurun_out/lab_scalar_best.txt 2>/dev/null || echo none)
if [ "$best" != "none" ]; then
  cp lab_build/s/$best.so cuda-to-sycl-nbody_b200/lib/libnbody_b200.so
  sha256sum cuda-to-sycl-nbody_b200/lib/libnbody_b200.so > gpurun_out/lab6_lib_sha.txt
  timeout 170 python -m pytest tests -m gpu -x -q > gpurun_out/lab6_pytest_gpu.txt 2>&1; tail -n 3 gpurun_out/lab6_pytest_gpu.txt
  timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/lab6_smoke.txt 2>&1; tail -n 3 gpurun_out/lab6_smoke.txt
  timeout 60 python tools/small_n.py > gpurun_out/lab6_small_n.txt 2>&1; cat gpurun_out/lab6_small_n.txt | cut -c1-200
fi
