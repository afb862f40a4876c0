#!/bin/bash
# usage: tools/ncu_one.sh <kernel regex> <tag> <skip> <command...>: one full ncu capture, summaries written on the box
mkdir -p gpurun_out
O=gpurun_out
K=$1; TAG=$2; SKIP=$3; shift 3
rm -f $O/$TAG.ncu-rep
ncu --set full --import-source on --clock-control none -k regex:$K -s $SKIP -c 1 -o $O/$TAG "$@" > $O/ncu_$TAG.log 2>&1
python profiles/ncu_summary.py $O/$TAG.ncu-rep --stalls > $O/${TAG}_ncu_summary.txt 2>&1
python profiles/ncu_stalls.py $O/$TAG.ncu-rep $K 2>/dev/null | head -60 > $O/${TAG}_stalls.txt
python profiles/ncu_linesamples.py $O/$TAG.ncu-rep 60 > $O/${TAG}_lines.txt 2>&1
[ "$KEEP_REP" = 1 ] || rm -f $O/$TAG.ncu-rep
grep -E "time_duration|inst_executed.sum|issue_active|registers" $O/${TAG}_ncu_summary.txt
