#!/bin/bash
# 4-GPU job: N = 1, 2, 4 pretrain bench back to back on one box (the driver's scaling sequence, without N = 8)
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
nvidia-smi -L > $O/n4f_gpus.txt
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-block-bench > $O/n4f_bench_n1.log 2>&1; echo "n1 rc=$?"
for N in 2 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-block-bench > $O/n4f_bench_n$N.log 2>&1; echo "n$N rc=$?"
done
