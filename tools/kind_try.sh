#!/bin/bash
# 4-state kernel variants on one box: monolithic vs specialised (KIND) launches, CTA sizes.  -> gpurun_out/kind_try.txt
mkdir -p gpurun_out
V=$PWD/bpp_b200/variants
run() { echo -n "$1: "; shift; env "$@" timeout 300 python tools/device_time.py $CFG $N $SC 2>&1 | tail -1; }
{
for rep in 1 2; do
CFG=config3 N=10000 SC=0
run "c3 mono      " BPPGPU_S4_KINDS=0
run "c3 kinds     " BPPGPU_S4_KINDS=1
run "c3 kinds cpt2" BPPGPU_S4_KINDS=1 BPPGPU_CPT=2
run "c3 nt512 cpt2 kinds" BPPGPU_LIB=$V/libbppgpu_nt512.so BPPGPU_CPT=2
run "c3 nt512 cpt2 mono " BPPGPU_LIB=$V/libbppgpu_nt512.so BPPGPU_CPT=2 BPPGPU_S4_KINDS=0
done
CFG=config3 N=10000 SC=1
run "c3s mono     " BPPGPU_S4_KINDS=0
run "c3s kinds    " BPPGPU_S4_KINDS=1
run "c3s nt512 cpt2 kinds" BPPGPU_LIB=$V/libbppgpu_nt512.so BPPGPU_CPT=2
CFG=config2 N=10000 SC=0
run "c2 mono      " BPPGPU_S4_KINDS=0
run "c2 kinds     " BPPGPU_S4_KINDS=1
run "c2 kinds cpt4" BPPGPU_S4_KINDS=1 BPPGPU_CPT=4
run "c2 nt512 cpt1 kinds" BPPGPU_LIB=$V/libbppgpu_nt512.so BPPGPU_CPT=1
CFG=config2 N=10000 SC=1
run "c2s mono     " BPPGPU_S4_KINDS=0
run "c2s kinds    " BPPGPU_S4_KINDS=1
} > gpurun_out/kind_try.txt 2>&1
cat gpurun_out/kind_try.txt
