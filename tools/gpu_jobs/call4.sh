#!/bin/bash
# GPU job 4: tests + smoke, 16-warp GEMM epilogues, row kernels, merged vs unmerged bench, launch list of the merged step
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r4_tests.log 2>&1; echo "tests rc=$?"; grep -E "^FAILED|passed|failed" $O/r4_tests.log | tail -12
MOME_BUILD_CACHED=1 python __graft_entry__.py smoke > $O/r4_smoke.log 2>&1; echo "smoke rc=$?"; grep "smoke" $O/r4_smoke.log
python tools/gemm_bench.py > $O/r4_gb.log 2>&1; cat $O/r4_gb.log
python tools/row_bench.py > $O/r4_row.log 2>&1; cat $O/r4_row.log
python bench.py --steps 10 --warmup 3 > $O/r4_bench_merged.log 2>&1; tail -c 1500 $O/r4_bench_merged.log
python bench.py --steps 10 --warmup 3 --no-merge --no-cpu-baseline --no-block-bench > $O/r4_bench_nomerge.log 2>&1; tail -c 300 $O/r4_bench_nomerge.log
export MOME_ATTN_TC=1 MOME_ATTN_TC_BWD=1
python bench.py --ncu-step > $O/r4_plain.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r4_launches.csv python bench.py --ncu-step > $O/r4_ncu.log 2>&1
tail -2 $O/r4_ncu.log
