#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" > $O/r12_tests.log 2>&1; echo "tests rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|timed out|Error|assert" $O/r12_tests.log | tail -8
timeout 300 python tools/attn_bench.py --check --tc-bwd p --iters 10 --long > $O/r12_attn_long.log 2>&1; echo "attn long rc=$?"; tail -5 $O/r12_attn_long.log | cut -c1-200
