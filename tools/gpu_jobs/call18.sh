#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
timeout 300 python tools/attn_bench.py --check --tc-bwd p --iters 20 --no-bwd > $O/r18_attn.log 2>&1; echo "rc=$?"; tail -9 $O/r18_attn.log | cut -c1-120
