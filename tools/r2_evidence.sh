#!/bin/bash
# Round-2 evidence pass on one B200: tests, bench lines, ncu launch list and full captures, tools.  Outputs -> gpurun_out/
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/*.ncu-rep
python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/r2_gpu_tests.log; cat $O/r2_gpu_tests.log
python bench.py --steps 100 --warmup 5 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; tail -c 600 $O/r2_bench_n1.json
python bench.py --impl reference --steps 10 --warmup 2 > $O/r2_bench_reference.json 2>> $O/r2_bench_n1.err
python bench.py --config config3 --scaling 1 --steps 50 --no-sub > $O/r2_bench_c3_scaling.json 2>> $O/r2_bench_n1.err
# launch list of the bench command (kernel shares of a step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_config3.csv python bench.py --steps 2 --warmup 3 --no-sub --no-cpu-baseline > $O/ncu_launches.log 2>&1
# full captures of the dominant kernels at full size; summarised on the box (gpurun_out/ is limited to 64 MiB)
summ() {   # $1 = report base name, $2 = kernel name fragment
  python profiles/ncu_summary.py $O/$1.ncu-rep --stalls > $O/$1_ncu_summary.txt 2>&1
  python profiles/ncu_stalls.py $O/$1.ncu-rep $2 2>/dev/null | head -45 > $O/$1_stalls.txt
  [ "$3" = keep ] || rm -f $O/$1.ncu-rep
}
ncu --set full --import-source on --clock-control none -k regex:tree_kernel_s4 -s 3 -c 1 -o $O/r2_tree_config3 python tools/device_time.py config3 10000 0 > $O/ncu_c3.log 2>&1; summ r2_tree_config3 tree_kernel_s4 keep
ncu --set full --import-source on --clock-control none -k regex:tree_kernel_s4 -s 3 -c 1 -o $O/r2_tree_config5 python tools/device_time.py config5 6250 0 > $O/ncu_c5.log 2>&1; summ r2_tree_config5 tree_kernel_s4
ncu --set full --import-source on --clock-control none -k regex:tree_kernel_s4 -s 3 -c 1 -o $O/r2_tree_config2 python tools/device_time.py config2 10000 0 > $O/ncu_c2.log 2>&1; summ r2_tree_config2 tree_kernel_s4
ncu --set full --import-source on --clock-control none -k regex:tree_kernel_s20c -s 3 -c 1 -o $O/r2_tree_config4 python tools/device_time.py config4 2000 0 > $O/ncu_c4.log 2>&1; summ r2_tree_config4 tree_kernel_s20c
ncu --set full --import-source on --clock-control none -k regex:plan_kernel_blocks -s 4 -c 1 -o $O/r2_plan_partial python tools/partial_update_bench.py 8 > $O/ncu_plan.log 2>&1; summ r2_plan_partial plan_kernel_blocks
python tools/partial_update_bench.py 8 16 48 > $O/r2_partial_update.jsonl 2> $O/r2_partial_update.err
python tools/tips_sweep.py 4 GTR > $O/r2_tips_sweep.txt 2>&1
python tools/tips_sweep.py 1 JC69 >> $O/r2_tips_sweep.txt 2>&1
rm -f $O/bpp_binary_report.txt
BPP_LONG=1 python -m pytest tests/test_bpp_binary.py -m gpu -q 2>&1 | tail -4 > $O/r2_bpp_long.log; cat $O/r2_bpp_long.log
cp $O/bpp_binary_report.txt $O/r2_bpp_binary_report.txt
ls -la $O | tail -30
