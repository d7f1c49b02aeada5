#!/bin/bash
# Round-2 sweep 5 on ONE B200 (under gpurun): generated tile body of the SCALAR one-body-per-lane kernel (small shards),
# variants of its period template, against ptxas' schedule (V0); plus R = 2 / R = 4 at the same sizes for the AUTO cost model.
mkdir -p gpurun_out
run() {  # lib, bodies, cfg
  export NBODY_KERNEL_CONFIG="$3"
  printf "%-6s N=%-7s cfg=%-7s " $(basename $1 .so) $2 "$3"
  NBODY_LAB_PARITY=1 timeout 60 python tools/lab_one.py $1 $2 8 || echo FAILED
}
{
  for n in 12800 25600 40000 51200 57720 80000; do
    for f in lab_build/s/V0.so lab_build/s/S?.so; do run $f $n 1,32,6; done
    run lab_build/s/V0.so $n 2,32,4; run lab_build/s/V0.so $n 4,32,4
  done
  for n in 6400 20000 32768 65536 100000; do run lab_build/s/V0.so $n 1,32,6; run lab_build/s/Sa.so $n 1,32,6; run lab_build/s/Sb.so $n 1,32,6; run lab_build/s/V0.so $n 2,32,4; run lab_build/s/V0.so $n 4,32,4; done
} > gpurun_out/lab5.txt 2>&1
cat gpurun_out/lab5.txt | cut -c1-150
