#!/bin/bash
python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -2
for n in 1048576 524288; do
python tools/run_steps.py --n $n --kernel auto --cfg 4,32,3 --steps 3 | tail -1 | cut -c12-200
for sg in 1 2 4 8 16; do NBODY_SEGS=$sg python tools/run_steps.py --n $n --kernel auto --cfg 4,32,4 --steps 3 | tail -1 | cut -c12-200 | sed "s/^/S=$sg /"; done
done
n=262144
python tools/run_steps.py --n $n --kernel auto --cfg 2,32,3 --steps 3 --iters 8 | tail -1 | cut -c12-200
for sg in 2 4 8 16; do NBODY_SEGS=$sg python tools/run_steps.py --n $n --kernel auto --cfg 2,32,4 --steps 3 --iters 8 | tail -1 | cut -c12-200 | sed "s/^/S=$sg /"; done
for sg in 2 4 8 16; do NBODY_SEGS=$sg python tools/run_steps.py --n $n --kernel auto --cfg 4,32,4 --steps 3 --iters 8 | tail -1 | cut -c12-200 | sed "s/^/S=$sg /"; done
