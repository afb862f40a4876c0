#!/bin/bash
# Closing run of round 2 after the big-tree changes: tests, smoke, bench line, tips sweeps, partial updates, sanitizer.
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/r2f_gpu_tests.log; cat $O/r2f_gpu_tests.log
grep -q " failed" $O/r2f_gpu_tests.log && exit 1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 100 --warmup 5 > $O/r2f_bench_n1.json 2> $O/r2f_bench_n1.err; tail -c 300 $O/r2f_bench_n1.json
python tools/tips_sweep.py 4 GTR > $O/r2f_tips_sweep_r4.txt 2>&1
python tools/tips_sweep.py 1 JC69 > $O/r2f_tips_sweep_r1.txt 2>&1
cut -c1-120 $O/r2f_tips_sweep_r4.txt $O/r2f_tips_sweep_r1.txt
tools/sanitize.sh
