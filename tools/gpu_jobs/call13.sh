#!/bin/bash
# GPU job 13: full GPU test suite + smoke, pretrain / vqa480 bench with the pipelined attention backward
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r13_tests.log 2>&1; echo "tests rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" $O/r13_tests.log | tail -12
MOME_BUILD_CACHED=1 python __graft_entry__.py smoke > $O/r13_smoke.log 2>&1; echo "smoke rc=$?"; grep "smoke" $O/r13_smoke.log
python bench.py --steps 10 --warmup 3 > $O/r13_bench.log 2>&1; echo "bench rc=$?"; tail -c 1500 $O/r13_bench.log
python bench.py --workload vqa480 --steps 8 --warmup 3 --no-cpu-baseline > $O/r13_bench_vqa480.log 2>&1; echo "bench vqa rc=$?"; tail -c 600 $O/r13_bench_vqa480.log
