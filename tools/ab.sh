#!/bin/bash
# A/B of two library builds on the same box, alternating: tools/ab.sh <variant> <config> <n> <scaling>
mkdir -p gpurun_out
for i in 1 2 3; do
  for v in "" $1; do
    if [ -n "$v" ]; then export BPPGPU_LIB=$PWD/bpp_b200/variants/libbppgpu_$v.so; else unset BPPGPU_LIB; fi
    echo -n "${v:-base}: "; timeout 300 python tools/device_time.py $2 $3 $4 2>&1 | tail -1
  done
done > gpurun_out/ab.txt 2>&1
cat gpurun_out/ab.txt
