#!/bin/bash
n=1048576
for mb in 20 24 25 26 28 30 32; do NBODY_MINB=$mb NBODY_SEGS=16 python tools/run_steps.py --n $n --kernel auto --cfg 4,32,4 --steps 3 | tail -1 | cut -c12-200 | sed "s/^/MINB=$mb /"; done
for sg in 32 64 128; do NBODY_SEGS=$sg python tools/run_steps.py --n $n --kernel auto --cfg 4,32,4 --steps 3 | tail -1 | cut -c12-200 | sed "s/^/S=$sg /"; done
for sg in 16 64; do NBODY_SEGS=$sg python tools/run_steps.py --n $n --kernel auto --cfg 2,32,4 --steps 3 | tail -1 | cut -c12-200 | sed "s/^/S=$sg /"; done
n=262144
for sg in 32 64; do NBODY_SEGS=$sg python tools/run_steps.py --n $n --kernel auto --cfg 2,32,4 --steps 3 --iters 8 | tail -1 | cut -c12-200 | sed "s/^/S=$sg /"; done
