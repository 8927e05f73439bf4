#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
MOME_ATTN_TC_MIN=16 timeout 300 python tools/attn_bench.py --check --tc-bwd p --iters 20 --only text > $O/r18_attn_text.log 2>&1; echo "rc=$?"; tail -3 $O/r18_attn_text.log | cut -c1-200
MOME_ATTN_TC_MIN=16 timeout 300 python tools/attn_bench.py --check --tc-bwd p --iters 20 --only text --batch 1024 > $O/r18_attn_text_1024.log 2>&1; echo "rc=$?"; tail -3 $O/r18_attn_text_1024.log | cut -c1-200
