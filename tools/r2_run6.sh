#!/bin/bash
mkdir -p gpurun_out
{
echo "== default"
python tools/device_time.py config3 10000 0
python tools/device_time.py config2 10000 0
export BPPGPU_LIB=$PWD/bpp_b200/variants/libbppgpu_nt512.so
echo "== nt512 cpt2"
BPPGPU_CPT=2 python tools/device_time.py config3 10000 0
BPPGPU_CPT=2 python tools/device_time.py config3 10000 1
echo "== nt512 cpt1 / cpt2 config2"
BPPGPU_CPT=1 python tools/device_time.py config2 10000 0
BPPGPU_CPT=2 python tools/device_time.py config2 10000 0
echo "== nt512 cpt4 config3"
BPPGPU_CPT=4 python tools/device_time.py config3 10000 0
} > gpurun_out/r2_devtime6.txt 2>&1
cat gpurun_out/r2_devtime6.txt
