#!/bin/bash
# 8-GPU job (final): pretrain step (fp32 and bf16 gradient reduction), ITC workload
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
nvidia-smi -L > $O/n8f_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 8 --steps 10 --warmup 3 > $O/n8f_bench.log 2>&1; echo "rc=$?"; tail -c 300 $O/n8f_bench.log
$TR bench.py --gpus 8 --steps 10 --warmup 3 --reduce-dtype bf16 --no-cpu-baseline --no-block-bench > $O/n8f_bench_bf16.log 2>&1; echo "rc=$?"; tail -c 300 $O/n8f_bench_bf16.log
$TR bench.py --gpus 8 --steps 8 --warmup 3 --workload itc4096 --no-cpu-baseline --no-block-bench > $O/n8f_bench_itc.log 2>&1; echo "rc=$?"; tail -c 300 $O/n8f_bench_itc.log
