#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-block-bench > $O/r23_bench.log 2>&1; echo "bench rc=$?"; tail -c 300 $O/r23_bench.log
MOME_BWD_SIDE_STREAM=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-block-bench > $O/r23_bench_off.log 2>&1; echo "bench off rc=$?"; tail -c 300 $O/r23_bench_off.log
python -m pytest tests -x -q -m gpu > $O/r23_tests.log 2>&1; echo "tests rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" $O/r23_tests.log | tail -5
