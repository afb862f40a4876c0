#!/bin/bash
mkdir -p gpurun_out
for c in 4 2 1; do echo "CPT=$c"; BPPGPU_CPT=$c timeout 300 python tools/device_time.py config3 10000 0 2>&1 | tail -1; done > gpurun_out/cpt_try.txt 2>&1
cat gpurun_out/cpt_try.txt
