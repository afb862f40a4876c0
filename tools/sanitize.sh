#!/bin/bash
# compute-sanitizer over the GPU parity suite (the BPP-binary tests run the reference's own process and are left out)
mkdir -p gpurun_out
O=gpurun_out
SEL="tests/test_gpu_parity.py tests/test_host_c.py"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 0 python -m pytest $SEL -m gpu -q -x 2>&1 | tail -8 > $O/r2f_san_mem.log; cat $O/r2f_san_mem.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 0 python -m pytest $SEL -m gpu -q -x 2>&1 | tail -8 > $O/r2f_san_race.log; cat $O/r2f_san_race.log
