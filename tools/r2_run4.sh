#!/bin/bash
mkdir -p gpurun_out
{
for v in "" noroot notma old; do
  if [ -z "$v" ]; then echo "== default"; unset BPPGPU_LIB; else echo "== $v"; export BPPGPU_LIB=$PWD/bpp_b200/variants/libbppgpu_$v.so; fi
  python tools/device_time.py config3 10000 0
  python tools/device_time.py config2 10000 0
done
} > gpurun_out/r2_devtime4.txt 2>&1
cat gpurun_out/r2_devtime4.txt
