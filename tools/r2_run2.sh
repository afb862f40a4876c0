#!/bin/bash
# GPU run 2 of round 2: parity after the lane permutation, store-path microbenchmark, bench, ncu
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_tests2.log
tail -3 gpurun_out/r2_tests2.log
./profiles/microbench/store_paths > gpurun_out/r2_store_paths.txt 2>&1
for cpt in 4 2; do BPPGPU_CPT=$cpt python tools/device_time.py config3 10000 0; done > gpurun_out/r2_devtime2.txt 2>&1
python tools/device_time.py config3 10000 1 >> gpurun_out/r2_devtime2.txt 2>&1
python tools/device_time.py config2 10000 0 >> gpurun_out/r2_devtime2.txt 2>&1
python tools/device_time.py config2 10000 1 >> gpurun_out/r2_devtime2.txt 2>&1
cat gpurun_out/r2_devtime2.txt
python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
python tools/partial_update_bench.py 8 16 48 > gpurun_out/r2_partial_update.jsonl 2> gpurun_out/r2_partial_update.err
ncu --set full --import-source on --clock-control none -k regex:tree_kernel_s4 -s 3 -c 1 -o gpurun_out/r2a_config3 python tools/device_time.py config3 4000 0 > gpurun_out/ncu_r2a.log 2>&1
tail -2 gpurun_out/ncu_r2a.log
