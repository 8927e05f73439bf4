#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
python -m pytest tests -x -q -m gpu > $O/r22_tests.log 2>&1; echo "tests rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" $O/r22_tests.log | tail -5
python -c "import __graft_entry__ as g; g.smoke()" > $O/r22_smoke.log 2>&1; echo "smoke rc=$?"; grep "smoke" $O/r22_smoke.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-block-bench > $O/r22_bench.log 2>&1; echo "bench rc=$?"; tail -c 300 $O/r22_bench.log
MOME_ATTN_FUSED_BIAS=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-block-bench > $O/r22_bench_unfused.log 2>&1; echo "bench rc=$?"; tail -c 300 $O/r22_bench_unfused.log
python tools/attn_bench.py --check --tc-bwd p --iters 20 --only fused > $O/r22_attn.log 2>&1; tail -3 $O/r22_attn.log | cut -c1-120
