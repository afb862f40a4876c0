#!/bin/bash
# multi-GPU checks: collective through the C-ABI from C (two engines in one process) and bench.py under torchrun
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_comm.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r2_comm_tests_n$N.log
cat gpurun_out/r2_comm_tests_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -c 1500 gpurun_out/r2_bench_n$N.json; tail -5 gpurun_out/r2_bench_n$N.err
