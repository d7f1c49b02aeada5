#!/bin/bash
# Round-2 sweep on ONE B200 (under gpurun): generated tile bodies (tools/sass_gen.py) vs the post-scheduled ptxas
# code (tools/sass_sched.py), every run parity-checked (SHA-256 of the forces; golden where one exists).
mkdir -p gpurun_out
run() {  # lib, bodies, cfg ("" = AUTO), steps
  if [ -n "$3" ]; then export NBODY_KERNEL_CONFIG="$3"; else unset NBODY_KERNEL_CONFIG; fi
  printf "%-28s N=%-8s cfg=%-7s " $(basename $1) $2 "${3:-auto}"
  NBODY_LAB_PARITY=1 timeout 120 python tools/lab_one.py lab_build/$1 $2 ${4:-3} || echo FAILED
}
{
  nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
  echo "## N = 1048576"
  run L1_prod_old.so 1048576 "" 4
  run L2_gen_old.so 1048576 6,32,4
  run L2_gen_old.so 1048576 8,32,4
  run L3_gen_lean.so 1048576 4,32,4
  run L3_gen_lean.so 1048576 6,32,4 4
  run L3_gen_lean.so 1048576 8,32,4 4
  run L4_sched_lean.so 1048576 6,32,4
  run L4_sched_lean.so 1048576 8,32,4
  run base_lean.so 1048576 8,32,4
  for l in L5a_split1 L5b_noqreuse L5c_gap10 L5d_nobetween; do run $l.so 1048576 6,32,4; run $l.so 1048576 8,32,4; done
  run L5e_r8_depth2.so 1048576 8,32,4
  run L5f_r8_depth2_split1.so 1048576 8,32,4
  run L7_gen_w8.so 1048576 6,32,4
  run L7_gen_w4.so 1048576 6,32,4
  echo "## N = 262144"
  run L1_prod_old.so 262144 "" 5
  for r in 2 4 6 8; do run L3_gen_lean.so 262144 $r,32,4 5; done
  run L4_sched_lean.so 262144 2,32,4 5
  run L4_sched_lean.so 262144 6,32,4 5
  echo "## N = 131072"
  run L1_prod_old.so 131072 "" 5
  for r in 2 4 6; do run L3_gen_lean.so 131072 $r,32,4 5; done
  echo "## N = 65536"
  run L1_prod_old.so 65536 "" 5
  for r in 2 4; do run L3_gen_lean.so 65536 $r,32,4 5; done
  echo "## N = 400003 (ragged)"
  run L1_prod_old.so 400003 "" 3
  run L3_gen_lean.so 400003 6,32,4 3
  run L3_gen_lean.so 400003 8,32,4 3
} > gpurun_out/lab2.txt 2>&1
cat gpurun_out/lab2.txt
