#!/bin/bash
for n in 1048576 524288; do
python tools/run_steps.py --n $n --kernel auto --cfg 4,32,3 --steps 3 | tail -1 | cut -c12-200
for t in 1 2 3 4 6 8 12; do python tools/run_steps.py --n $n --kernel auto --cfg $t,32,4 --steps 3 | tail -1 | cut -c12-200 | sed "s/^/T=$t /"; done
done
for n in 262144 131072; do
python tools/run_steps.py --n $n --kernel auto --cfg 2,32,3 --steps 3 --iters 8 | tail -1 | cut -c12-200
for t in 2 4 8; do python tools/run_steps.py --n $n --kernel auto --cfg $t,32,4 --steps 3 --iters 8| tail -1 | cut -c12-200 | sed "s/^/T=$t /"; done
done
