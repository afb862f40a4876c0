#!/bin/bash
# 4-state kernel: monolithic vs specialised (KIND) launches.  -> gpurun_out/kind_try.txt
mkdir -p gpurun_out
run() { echo -n "$1: "; shift; env "$@" timeout 300 python tools/device_time.py $CFG $N $SC 2>&1 | tail -1; }
{
for rep in 1 2; do
CFG=config3 N=10000 SC=0
run "c3 mono " BPPGPU_S4_KINDS=0
run "c3 kinds" BPPGPU_S4_KINDS=1
CFG=config3 N=10000 SC=1
run "c3s mono " BPPGPU_S4_KINDS=0
run "c3s kinds" BPPGPU_S4_KINDS=1
CFG=config2 N=10000 SC=1
run "c2s mono " BPPGPU_S4_KINDS=0
run "c2s kinds" BPPGPU_S4_KINDS=1
done
CFG=config2 N=10000 SC=0
run "c2 mono " BPPGPU_S4_KINDS=0
run "c2 kinds" BPPGPU_S4_KINDS=1
} > gpurun_out/kind_try.txt 2>&1
cat gpurun_out/kind_try.txt
