#!/bin/bash
# 20-state kernel check: parity tests that touch 20 states, device time of config 4 (unscaled / scaled)
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_bpp_binary.py 2>&1 | tail -5 > $O/s20_tests.log; cat $O/s20_tests.log
for s in 0 1; do timeout 300 python tools/device_time.py config4 2000 $s; done > $O/s20_devtime.txt 2>&1; cat $O/s20_devtime.txt
BPPGPU_S20_FUSE_IMAGES=0 timeout 300 python tools/device_time.py config4 2000 0 2>&1 | tail -1
