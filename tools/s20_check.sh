#!/bin/bash
# 20-state kernel check: parity tests that touch 20 states, device time of config 4 (unscaled / scaled)
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "lg or LG or 20 or partial or split or ragged or dense or vector or golden or fixture or shape or model" 2>&1 | tail -15 > $O/s20_tests.log; cat $O/s20_tests.log
for s in 0 1; do timeout 300 python tools/device_time.py config4 2000 $s; done > $O/s20_devtime.txt 2>&1; cat $O/s20_devtime.txt
