#!/bin/bash
# Last run of round 2: tests, smoke, bench line, category-count sweeps, sanitizer.
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/r2f_gpu_tests.log; cat $O/r2f_gpu_tests.log
grep -q " failed" $O/r2f_gpu_tests.log && exit 1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 100 --warmup 5 > $O/r2f_bench_n1.json 2> $O/r2f_bench_n1.err; tail -c 300 $O/r2f_bench_n1.json
{ python tools/tips_sweep.py 3 GTR 8,16,32; python tools/tips_sweep.py 5 GTR 8,16,32,64; python tools/tips_sweep.py 8 GTR 8,16,32,64; python tools/tips_sweep.py 2 GTR 8,16,32,64; } > $O/r2f_tips_sweep_cats.txt 2>&1
cut -c1-120 $O/r2f_tips_sweep_cats.txt
tools/sanitize.sh
