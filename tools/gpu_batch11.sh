#!/bin/bash
for n in 524288 262144; do
for sg in 16 64; do
export NBODY_SEGS=$sg
NBODY_MINB=14 python tools/run_steps.py --n $n --kernel auto --cfg 6,32,4 --steps 3 --iters 4 | tail -1 | cut -c12-200 | sed "s/^/S=$sg R6 MINB=14 /"
NBODY_MINB=20 python tools/run_steps.py --n $n --kernel auto --cfg 4,32,4 --steps 3 --iters 4 | tail -1 | cut -c12-200 | sed "s/^/S=$sg R4 MINB=20 /"
NBODY_MINB=28 python tools/run_steps.py --n $n --kernel auto --cfg 4,32,4 --steps 3 --iters 4 | tail -1 | cut -c12-200 | sed "s/^/S=$sg R4 MINB=28 /"
NBODY_MINB=28 python tools/run_steps.py --n $n --kernel auto --cfg 2,32,4 --steps 3 --iters 4 | tail -1 | cut -c12-200 | sed "s/^/S=$sg R2 MINB=28 /"
NBODY_MINB=20 python tools/run_steps.py --n $n --kernel auto --cfg 2,32,4 --steps 3 --iters 4 | tail -1 | cut -c12-200 | sed "s/^/S=$sg R2 MINB=20 /"
done; done
n=4194304
NBODY_SEGS=8 NBODY_MINB=14 python tools/run_steps.py --n $n --kernel auto --cfg 6,32,4 --steps 2 | tail -1 | cut -c12-200 | sed "s/^/4M S=8 R6 MINB=14 /"
NBODY_SEGS=1 NBODY_MINB=14 python tools/run_steps.py --n $n --kernel auto --cfg 6,32,4 --steps 2 | tail -1 | cut -c12-200 | sed "s/^/4M S=1 R6 MINB=14 /"
