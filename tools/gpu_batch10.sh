#!/bin/bash
n=1048576
export NBODY_SEGS=32
for mb in 14 16 18 20 22; do NBODY_MINB=$mb python tools/run_steps.py --n $n --kernel auto --cfg 4,32,4 --steps 3 | tail -1 | cut -c12-200 | sed "s/^/R4 MINB=$mb /"; done
for mb in 12 14 16 18 20; do NBODY_MINB=$mb python tools/run_steps.py --n $n --kernel auto --cfg 6,32,4 --steps 3 | tail -1 | cut -c12-200 | sed "s/^/R6 MINB=$mb /"; done
for mb in 10 12 14 16; do NBODY_MINB=$mb python tools/run_steps.py --n $n --kernel auto --cfg 8,32,4 --steps 3 | tail -1 | cut -c12-200 | sed "s/^/R8 MINB=$mb /"; done
for mb in 20 28 32; do NBODY_MINB=$mb python tools/run_steps.py --n $n --kernel auto --cfg 2,32,4 --steps 3 | tail -1 | cut -c12-200 | sed "s/^/R2 MINB=$mb /"; done
