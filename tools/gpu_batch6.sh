#!/bin/bash
n=1048576
for cfg in 4,16,5 4,20,5 4,24,5 4,28,5 4,32,5 6,16,5 6,20,5 8,12,5 8,16,5 2,24,5 2,32,5; do python tools/run_steps.py --n $n --kernel auto --cfg $cfg --steps 3 | tail -1 | cut -c12-200 | sed "s/^/cfg=$cfg /"; done
for t in 1 2 3 4 6 8; do python tools/run_steps.py --n $n --kernel auto --cfg $t,32,4 --steps 3 | tail -1 | cut -c12-200 | sed "s/^/T=$t /"; done
n=524288
for t in 1 2 3 4 6 8; do python tools/run_steps.py --n $n --kernel auto --cfg $t,32,4 --steps 3 | tail -1 | cut -c12-200 | sed "s/^/T=$t /"; done
for cfg in 4,24,5 4,28,5 2,24,5; do python tools/run_steps.py --n $n --kernel auto --cfg $cfg --steps 3 | tail -1 | cut -c12-200 | sed "s/^/cfg=$cfg /"; done
