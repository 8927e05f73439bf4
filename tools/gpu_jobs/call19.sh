#!/bin/bash
# final check of HEAD: the driver's sequence (GPU tests, smoke, default bench, reference arm)
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
python -m pytest tests -x -q -m gpu > $O/r19_tests.log 2>&1; echo "tests rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" $O/r19_tests.log | tail -5
python -c "import __graft_entry__ as g; g.smoke()" > $O/r19_smoke.log 2>&1; echo "smoke rc=$?"; grep "smoke" $O/r19_smoke.log
python bench.py --impl reference > $O/r19_bench_ref.log 2>&1; echo "ref rc=$?"; tail -c 200 $O/r19_bench_ref.log
python bench.py > $O/r19_bench.log 2>&1; echo "bench rc=$?"; tail -c 300 $O/r19_bench.log
