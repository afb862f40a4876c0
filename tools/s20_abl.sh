#!/bin/bash
mkdir -p gpurun_out
for v in ng1 ng1b; do
  if [ -n "$v" ]; then export BPPGPU_LIB=$PWD/bpp_b200/variants/libbppgpu_$v.so; fi
  echo "== ${v:-base}"; timeout 300 python tools/device_time.py config4 2000 0 2>&1 | tail -1; timeout 300 python tools/device_time.py config4 2000 1 2>&1 | tail -1
done > gpurun_out/s20_abl.txt 2>&1
cat gpurun_out/s20_abl.txt
