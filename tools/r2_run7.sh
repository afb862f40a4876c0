#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2_tests7.log
tail -3 gpurun_out/r2_tests7.log
{
python tools/device_time.py config3 10000 0
python tools/device_time.py config3 10000 1
python tools/device_time.py config2 10000 0
python tools/device_time.py config2 10000 1
python tools/device_time.py config5 6250 0
} > gpurun_out/r2_devtime7.txt 2>&1
cat gpurun_out/r2_devtime7.txt
ncu --set full --import-source on --clock-control none -k regex:tree_kernel_s4 -s 3 -c 1 -o gpurun_out/r2c_config3 python tools/device_time.py config3 4000 0 > gpurun_out/ncu_r2c.log 2>&1
tail -2 gpurun_out/ncu_r2c.log
