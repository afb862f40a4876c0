#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2_tests3.log
tail -3 gpurun_out/r2_tests3.log
{
echo "== default build (perm off, NSTAGE 2, TMA block load, one log per site)"
python tools/device_time.py config3 10000 0
python tools/device_time.py config3 10000 1
echo "== NSTAGE4=1"
BPPGPU_LIB=$PWD/bpp_b200/variants/libbppgpu_ns1.so python tools/device_time.py config3 10000 0
echo "== plan cache off"
BPPGPU_PLAN_CACHE=0 python tools/device_time.py config3 10000 0
python tools/device_time.py config2 10000 0
} > gpurun_out/r2_devtime3.txt 2>&1
cat gpurun_out/r2_devtime3.txt
python tools/tips_sweep.py 4 GTR > gpurun_out/r2_tips_sweep_r4.txt 2>&1
python tools/tips_sweep.py 1 JC69 > gpurun_out/r2_tips_sweep_r1.txt 2>&1
cat gpurun_out/r2_tips_sweep_r4.txt gpurun_out/r2_tips_sweep_r1.txt
