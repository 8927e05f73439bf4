#!/bin/bash
# 2-GPU job (final): whole GPU test suite incl. the 2-rank tests, N = 1 and N = 2 pretrain bench on the same box
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
nvidia-smi -L > $O/n2f_gpus.txt
python -m pytest tests -m gpu -q > $O/n2f_tests.log 2>&1; echo "tests rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" $O/n2f_tests.log | tail -8
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/n2f_bench_n1.log 2>&1; echo "n1 rc=$?"; tail -c 300 $O/n2f_bench_n1.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 10 --warmup 3 > $O/n2f_bench.log 2>&1; echo "n2 rc=$?"; tail -c 300 $O/n2f_bench.log
$TR bench.py --gpus 2 --steps 8 --warmup 3 --workload vqa480 --no-cpu-baseline > $O/n2f_bench_vqa480.log 2>&1; echo "n2 vqa rc=$?"; tail -c 300 $O/n2f_bench_vqa480.log
$TR bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > $O/n2f_bench_reference.log 2>&1; echo "n2 ref rc=$?"; tail -c 200 $O/n2f_bench_reference.log
