#!/bin/bash
# GPU job 9: full GPU test suite + smoke at HEAD, then ncu --set full of the tcgen05 attention backward / forward
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r9_tests.log 2>&1; echo "tests rc=$?"; grep -E "^FAILED|passed|failed" $O/r9_tests.log | tail -12
MOME_BUILD_CACHED=1 python __graft_entry__.py smoke > $O/r9_smoke.log 2>&1; echo "smoke rc=$?"; grep "smoke" $O/r9_smoke.log
python tools/attn_bench.py --only fused --tc-bwd 1 --iters 20 > $O/r9_attn.log 2>&1; echo "attn rc=$?"; tail -3 $O/r9_attn.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc_kernel -s 2 -c 1 -f -o $O/r9_attn_bwd \
  python tools/attn_bench.py --only fused --tc-bwd 1 --iters 1 > $O/r9_ncu_bwd.log 2>&1; echo "ncu bwd rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc_kernel -s 2 -c 1 -f -o $O/r9_attn_fwd \
  python tools/attn_bench.py --only fused --tc-bwd 1 --iters 1 > $O/r9_ncu_fwd.log 2>&1; echo "ncu fwd rc=$?"
