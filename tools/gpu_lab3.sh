#!/bin/bash
# Round-2 sweep 3 on ONE B200 (under gpurun): schedule templates / MUFU spacing / reuse flags of tools/sass_gen.py,
# all four register-blocking factors, every run parity-checked (SHA-256 of the forces; golden where one exists).
mkdir -p gpurun_out
run() {  # lib, bodies, cfg, steps
  export NBODY_KERNEL_CONFIG="$3"
  printf "%-22s N=%-8s cfg=%-7s " $(basename $1 .so) $2 "$3"
  NBODY_LAB_PARITY=1 timeout 120 python tools/lab_one.py $1 $2 ${4:-3} || echo FAILED
}
{
  nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
  echo "## R=6 N=1048576"; for f in lab_build/v/*.so; do run $f 1048576 6,32,4 3; done
  echo "## R=8 N=1048576"; for f in lab_build/v/*.so; do run $f 1048576 8,32,4 3; done
  echo "## R=4 N=262144"; for f in lab_build/v/*.so; do run $f 262144 4,32,4 5; done
  echo "## R=2 N=131072"; for f in lab_build/v/*.so; do run $f 131072 2,32,4 5; done
} > gpurun_out/lab3.txt 2>&1
grep -c parity=True gpurun_out/lab3.txt; grep -c -E "parity=False|FAILED" gpurun_out/lab3.txt
sort -t= -k5 gpurun_out/lab3.txt | grep "N=1048576" | sort -k6 | tail -5
