#!/bin/bash
# partial-update A/B on one box: default build against a variant.  usage: tools/pu_try.sh <variant>  -> gpurun_out/pu_try.txt
mkdir -p gpurun_out
V=$PWD/bpp_b200/variants/libbppgpu_$1.so
{
for rep in 1 2; do
  for v in base $1; do
    if [ $v = base ]; then unset BPPGPU_LIB; else export BPPGPU_LIB=$V; fi
    echo "== $v"; PU_NO_CPU=1 timeout 600 python tools/partial_update_bench.py ${TIPS:-8 16 48} 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print(d['tips'], 'dev ms %.3f' % d['ms_per_batch_device'], 'e2e us %.0f' % d['us_per_batch_e2e'], {k: round(v, 3) for k, v in d['per_kernel_ms'].items()}, d['kernel'], d.get('plan_stats'))
"
  done
done
} > gpurun_out/pu_try.txt 2>&1
cat gpurun_out/pu_try.txt
