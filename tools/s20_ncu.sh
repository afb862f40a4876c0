#!/bin/bash
# full ncu capture of the 20-state tree kernel at config-4 size; summaries written on the box
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/*.ncu-rep
SC=${1:-0}
TAG=${2:-s20t}
ncu --set full --import-source on --clock-control none -k regex:tree_kernel_s20 -s 3 -c 1 -o $O/$TAG python tools/device_time.py config4 2000 $SC > $O/ncu_$TAG.log 2>&1
python profiles/ncu_summary.py $O/$TAG.ncu-rep --stalls > $O/${TAG}_ncu_summary.txt 2>&1
python profiles/ncu_stalls.py $O/$TAG.ncu-rep tree_kernel_s20 2>/dev/null | head -60 > $O/${TAG}_stalls.txt
python profiles/ncu_linesamples.py $O/$TAG.ncu-rep 45 > $O/${TAG}_lines.txt 2>&1
python profiles/ncu_smem.py $O/$TAG.ncu-rep > $O/${TAG}_smem.txt 2>&1
cat $O/${TAG}_ncu_summary.txt
