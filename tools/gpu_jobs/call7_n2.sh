#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 8 --warmup 3 --no-overlap > $O/n2c_noov.log 2>&1; echo "rc=$?"; tail -c 300 $O/n2c_noov.log
$TR bench.py --gpus 2 --steps 8 --warmup 3 --no-overlap --reduce-dtype bf16 > $O/n2c_noov_bf16.log 2>&1; echo "rc=$?"; tail -c 300 $O/n2c_noov_bf16.log
$TR bench.py --gpus 2 --steps 8 --warmup 3 --no-overlap --zero2 > $O/n2c_noov_zero2.log 2>&1; echo "rc=$?"; tail -c 300 $O/n2c_noov_zero2.log
$TR bench.py --gpus 2 --steps 8 --warmup 3 > $O/n2c_ov.log 2>&1; echo "rc=$?"; tail -c 300 $O/n2c_ov.log
