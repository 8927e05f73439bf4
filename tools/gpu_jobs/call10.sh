#!/bin/bash
# GPU job 10: software-pipelined tcgen05 attention backward: tests, cross-check + timing against the first kernel
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" > $O/r10_tests.log 2>&1; echo "tests rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|timed out|Error" $O/r10_tests.log | tail -12
timeout 300 python tools/attn_bench.py --check --tc-bwd p --iters 20 > $O/r10_attn_p.log 2>&1; echo "attn p rc=$?"; tail -9 $O/r10_attn_p.log
timeout 300 python tools/attn_bench.py --tc-bwd 1 --iters 20 --only fused > $O/r10_attn_1.log 2>&1; echo "attn 1 rc=$?"; tail -3 $O/r10_attn_1.log
