#!/bin/bash
# 8-GPU job: pretrain step with gradient reduction overlapped / after the backward, ITC workload (BASELINE configs[1], configs[4] at full scale)
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
nvidia-smi -L > $O/n8_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 8 --steps 8 --warmup 3 > $O/n8_bench.log 2>&1; echo "rc=$?"; tail -c 300 $O/n8_bench.log
$TR bench.py --gpus 8 --steps 8 --warmup 3 --no-overlap > $O/n8_bench_noov.log 2>&1; echo "rc=$?"; tail -c 300 $O/n8_bench_noov.log
$TR bench.py --gpus 8 --steps 8 --warmup 3 --workload itc4096 > $O/n8_bench_itc.log 2>&1; echo "rc=$?"; tail -c 300 $O/n8_bench_itc.log
