#!/bin/bash
for cfg in 4,32,3 4,64,3 4,128,3 2,32,3 6,32,3; do python tools/run_steps.py --n 1048576 --kernel auto --cfg $cfg --steps 3 | tail -1 ; done
for cfg in 4,256,1 4,128,1 2,128,1; do python tools/run_steps.py --n 1048576 --kernel packed --cfg $cfg --steps 3 | tail -1 ; done
