#!/bin/bash
# 2-GPU job: distributed tests (ITC gather modes, GradSync full step), bench at N = 2 (all-reduce and ZeRO-2), ITC workload
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
nvidia-smi -L > $O/n2_gpus.txt
python -m pytest tests/test_dist_gpu.py -q > $O/n2_tests.log 2>&1; echo "dist tests rc=$?"; tail -5 $O/n2_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 8 --warmup 3 > $O/n2_bench.log 2>&1; echo "bench rc=$?"; tail -c 900 $O/n2_bench.log
$TR bench.py --gpus 2 --steps 8 --warmup 3 --zero2 > $O/n2_bench_zero2.log 2>&1; echo "zero2 rc=$?"; tail -c 600 $O/n2_bench_zero2.log
$TR bench.py --gpus 2 --steps 8 --warmup 3 --workload itc4096 > $O/n2_bench_itc.log 2>&1; echo "itc rc=$?"; tail -c 600 $O/n2_bench_itc.log
