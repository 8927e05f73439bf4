#!/bin/bash
# GPU job 5: bf16 transpose tile in the GEMM epilogues, 16-warp attention backward timing, merged bench, other workloads
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r5_tests.log 2>&1; echo "tests rc=$?"; grep -E "^FAILED|passed|failed" $O/r5_tests.log | tail -12
MOME_BUILD_CACHED=1 python __graft_entry__.py smoke > $O/r5_smoke.log 2>&1; echo "smoke rc=$?"; grep "smoke" $O/r5_smoke.log
python tools/gemm_bench.py > $O/r5_gb.log 2>&1; cat $O/r5_gb.log
python tools/attn_bench.py --check --tc-bwd 3 --iters 20 > $O/r5_attn_bwd3.log 2>&1; echo "attn tc-bwd=3 rc=$?"; tail -9 $O/r5_attn_bwd3.log
python bench.py --steps 10 --warmup 3 > $O/r5_bench.log 2>&1; tail -c 1200 $O/r5_bench.log
python bench.py --workload vqa480 --steps 8 --warmup 3 > $O/r5_bench_vqa480.log 2>&1; tail -c 300 $O/r5_bench_vqa480.log
python bench.py --workload itc4096 --steps 8 --warmup 3 > $O/r5_bench_itc4096.log 2>&1; tail -c 300 $O/r5_bench_itc4096.log
python bench.py --model vlmo_large --steps 6 --warmup 3 > $O/r5_bench_large.log 2>&1; tail -c 300 $O/r5_bench_large.log
python bench.py --steps 10 --warmup 3 --dedup --no-cpu-baseline --no-block-bench > $O/r5_bench_dedup.log 2>&1; tail -c 300 $O/r5_bench_dedup.log
