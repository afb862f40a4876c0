#!/bin/bash
# Final evidence pass of round 2 on one B200: tests, bench lines, ncu launch list and full captures of the kernels the
# bench times (the specialised 4-state launches, tree_kernel_s20t), tools.  Outputs -> gpurun_out/ (copied to profiles/)
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/*.ncu-rep
python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/r2f_gpu_tests.log; cat $O/r2f_gpu_tests.log
grep -q " failed" $O/r2f_gpu_tests.log && exit 1
python bench.py --steps 100 --warmup 5 > $O/r2f_bench_n1.json 2> $O/r2f_bench_n1.err; tail -c 400 $O/r2f_bench_n1.json
python bench.py --impl reference --steps 10 --warmup 2 > $O/r2f_bench_reference.json 2>> $O/r2f_bench_n1.err
python bench.py --config config3 --scaling 1 --steps 50 --no-sub > $O/r2f_bench_c3_scaling.json 2>> $O/r2f_bench_n1.err
python bench.py --config config4 --scaling 1 --steps 50 --no-sub --no-cpu-baseline > $O/r2f_bench_c4_scaling.json 2>> $O/r2f_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2f_launches_config3.csv python bench.py --steps 2 --warmup 3 --no-sub --no-cpu-baseline > $O/ncu_launches.log 2>&1
summ() {   # $1 = report base name, $2 = kernel name fragment
  python profiles/ncu_summary.py $O/$1.ncu-rep --stalls > $O/$1_ncu_summary.txt 2>&1
  python profiles/ncu_stalls.py $O/$1.ncu-rep $2 2>/dev/null | head -45 > $O/$1_stalls.txt
  rm -f $O/$1.ncu-rep
}
# device_time.py runs 5 warm-up steps (plan, classify, specialised launches from the third on) before the timed ones:
# -s 6 lands on a launch of the cached-plan steady state (the scaled-only instantiation for config 3 with scaling)
ncu --set full --import-source on --clock-control none -k regex:tree_kernel_s4 -s 6 -c 1 -o $O/r2f_tree_config3 python tools/device_time.py config3 10000 0 > $O/ncu_c3.log 2>&1; summ r2f_tree_config3 tree_kernel_s4
ncu --set full --import-source on --clock-control none -k regex:tree_kernel_s4 -s 6 -c 1 -o $O/r2f_tree_config3s python tools/device_time.py config3 10000 1 > $O/ncu_c3s.log 2>&1; summ r2f_tree_config3s tree_kernel_s4
ncu --set full --import-source on --clock-control none -k regex:tree_kernel_s4 -s 6 -c 1 -o $O/r2f_tree_config5 python tools/device_time.py config5 6250 0 > $O/ncu_c5.log 2>&1; summ r2f_tree_config5 tree_kernel_s4
ncu --set full --import-source on --clock-control none -k regex:tree_kernel_s4 -s 6 -c 1 -o $O/r2f_tree_config2 python tools/device_time.py config2 10000 0 > $O/ncu_c2.log 2>&1; summ r2f_tree_config2 tree_kernel_s4
ncu --set full --import-source on --clock-control none -k regex:tree_kernel_s20t -s 6 -c 1 -o $O/r2f_tree_config4 python tools/device_time.py config4 2000 0 > $O/ncu_c4.log 2>&1; summ r2f_tree_config4 tree_kernel_s20t
python tools/partial_update_bench.py 8 16 48 128 > $O/r2f_partial_update.jsonl 2> $O/r2f_partial_update.err
python tools/tips_sweep.py 4 GTR > $O/r2f_tips_sweep_r4.txt 2>&1
python tools/tips_sweep.py 1 JC69 > $O/r2f_tips_sweep_r1.txt 2>&1
ls -la $O | tail -30
