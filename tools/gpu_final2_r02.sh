#!/bin/bash
# Round-2 validation after the accumulator-relay kernel went into AUTO (ONE B200, ~1.5 GPU-minutes): the whole GPU
# test-suite, smoke(), the small-N table against the reference kernel, a quick bench line.
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest_gpu.txt 2>&1; tail -n 3 gpurun_out/r02b_pytest_gpu.txt
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke.txt 2>&1; tail -n 5 gpurun_out/r02b_smoke.txt | cut -c1-200
timeout 90 python tools/small_n.py 1024 2048 4096 6400 12800 14208 18944 25600 51200 > gpurun_out/r02b_small_n.txt 2>&1; cat gpurun_out/r02b_small_n.txt | cut -c1-220
