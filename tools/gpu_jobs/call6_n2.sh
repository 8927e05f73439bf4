#!/bin/bash
# 2-GPU job: gradient-reduction variants (overlap with NCCL capped / uncapped, after-backward fp32 / bf16)
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
python -m pytest tests/test_kernels_gpu.py -q -k "attention or attn" > $O/n2b_attn_tests.log 2>&1; echo "attn tests rc=$?"; tail -2 $O/n2b_attn_tests.log
python tools/attn_bench.py --tc-bwd 1 --iters 20 --no-bwd > $O/n2b_attn_fwd.log 2>&1; tail -8 $O/n2b_attn_fwd.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 8 --warmup 3 --no-overlap > $O/n2b_noov.log 2>&1; echo "rc=$?"; tail -c 400 $O/n2b_noov.log
NCCL_MAX_CTAS=32 $TR bench.py --gpus 2 --steps 8 --warmup 3 > $O/n2b_ov32.log 2>&1; echo "rc=$?"; tail -c 400 $O/n2b_ov32.log
$TR bench.py --gpus 2 --steps 8 --warmup 3 --no-overlap --reduce-dtype bf16 > $O/n2b_noov_bf16.log 2>&1; echo "rc=$?"; tail -c 400 $O/n2b_noov_bf16.log
$TR bench.py --gpus 2 --steps 8 --warmup 3 > $O/n2b_ov8.log 2>&1; echo "rc=$?"; tail -c 400 $O/n2b_ov8.log
