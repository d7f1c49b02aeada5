#!/bin/bash
for cfg in 4,32,3 ; do for k in 12 16 20 24 28 32; do python tools/run_steps.py --n 1048576 --kernel auto --cfg $cfg --resident $k --steps 2 | tail -1 | sed "s/^/k=$k /"; done; done
for cfg in 6,32,3 8,32,3; do for k in 8 12 16 18; do python tools/run_steps.py --n 1048576 --kernel auto --cfg $cfg --resident $k --steps 2 | tail -1 | sed "s/^/k=$k /"; done; done
for cfg in 2,32,3; do for k in 16 24 32; do python tools/run_steps.py --n 1048576 --kernel auto --cfg $cfg --resident $k --steps 2 | tail -1 | sed "s/^/k=$k /"; done; done
