#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" > $O/r12_tests.log 2>&1; echo "tests rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|timed out|Error" $O/r12_tests.log | tail -5
for v in p; do
timeout 300 python tools/attn_bench.py --check --tc-bwd $v --iters 20 > $O/r12_attn_$v.log 2>&1; echo "attn $v rc=$?"; tail -9 $O/r12_attn_$v.log | cut -c1-100
done
