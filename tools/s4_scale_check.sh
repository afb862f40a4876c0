#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_bpp_binary.py 2>&1 | tail -6 > gpurun_out/s4s_tests.log; cat gpurun_out/s4s_tests.log
for c in config3 config2; do for s in 0 1; do timeout 300 python tools/device_time.py $c 10000 $s 2>&1 | tail -1; done; done > gpurun_out/s4s_devtime.txt; cat gpurun_out/s4s_devtime.txt
