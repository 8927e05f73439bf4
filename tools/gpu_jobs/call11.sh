#!/bin/bash
# GPU job 11: ncu --set full of the pipelined attention backward
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc_pipe_kernel -s 2 -c 1 -f -o $O/r11_attn_bwd_pipe \
  python tools/attn_bench.py --only fused --tc-bwd p --iters 1 > $O/r11_ncu.log 2>&1; echo "ncu rc=$?"
