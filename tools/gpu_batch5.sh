#!/bin/bash
# mixed-grid tail sweep: cfg "T,32,4" = T 64-body units per SM sub-partition at the end of the queue
for n in 524288 1048576; do
python tools/run_steps.py --n $n --kernel auto --cfg 4,32,3 --steps 3 | tail -1 | cut -c1-200
for t in 2 4 8 12 16; do python tools/run_steps.py --n $n --kernel auto --cfg $t,32,4 --steps 3 | tail -1 | cut -c1-200 | sed "s/^/T=$t /"; done
done
python tools/run_steps.py --n 262144 --kernel auto --cfg 8,32,4 --steps 3 --iters 4| tail -1 | cut -c1-200
python tools/run_steps.py --n 262144 --kernel auto --cfg 2,32,3 --steps 3 --iters 4| tail -1 | cut -c1-200
python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -2
