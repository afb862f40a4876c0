#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2_tests5.log
tail -3 gpurun_out/r2_tests5.log
{
for v in "" notma; do
  if [ -z "$v" ]; then echo "== default"; unset BPPGPU_LIB; else echo "== $v"; export BPPGPU_LIB=$PWD/bpp_b200/variants/libbppgpu_$v.so; fi
  python tools/device_time.py config3 10000 0
  python tools/device_time.py config3 10000 1
  BPPGPU_CPT=2 python tools/device_time.py config3 10000 0
  python tools/device_time.py config5 6250 0
done
} > gpurun_out/r2_devtime5.txt 2>&1
cat gpurun_out/r2_devtime5.txt
unset BPPGPU_LIB
ncu --set full --import-source on --clock-control none -k regex:tree_kernel_s4 -s 3 -c 1 -o gpurun_out/r2b_config3 python tools/device_time.py config3 4000 0 > gpurun_out/ncu_r2b.log 2>&1
tail -2 gpurun_out/ncu_r2b.log
