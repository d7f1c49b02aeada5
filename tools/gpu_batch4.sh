#!/bin/bash
python tools/run_steps.py --n 524288 --kernel auto --cfg 4,32,3 --steps 3 --iters 8 | tail -1 | cut -c1-200
python tools/run_steps.py --n 524288 --kernel auto --cfg 2,32,3 --steps 3 --iters 8 | tail -1 | cut -c1-200
python tools/run_steps.py --n 1048576 --kernel auto --cfg 4,32,3 --steps 3 --iters 4 | tail -1 | cut -c1-200
python tools/run_steps.py --n 1048576 --kernel packed --cfg 4,256,1 --steps 3 --iters 4 | tail -1 | cut -c1-200
python tools/run_steps.py --n 262144 --kernel auto --steps 3 --iters 16 | tail -1 | cut -c1-200
