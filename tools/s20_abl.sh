#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "lg or LG or 20 or partial or split or ragged or dense or vector or golden or fixture or shape or model" 2>&1 | tail -15 > gpurun_out/s20_tests.log; cat gpurun_out/s20_tests.log
for v in "" abl_NODMMA abl_NOSTORE; do
  if [ -n "$v" ]; then export BPPGPU_LIB=$PWD/bpp_b200/variants/libbppgpu_$v.so; fi
  echo "== ${v:-base}"; timeout 300 python tools/device_time.py config4 2000 0 2>&1 | tail -1
done > gpurun_out/s20_abl.txt 2>&1
unset BPPGPU_LIB
timeout 300 python tools/device_time.py config4 2000 1 2>&1 | tail -1 >> gpurun_out/s20_abl.txt
cat gpurun_out/s20_abl.txt
