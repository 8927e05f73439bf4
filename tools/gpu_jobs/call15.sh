#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" > $O/r15_tests.log 2>&1; echo "tests rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|timed out|Error|assert" $O/r15_tests.log | tail -8
