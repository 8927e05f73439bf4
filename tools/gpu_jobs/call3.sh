#!/bin/bash
# GPU job 3: tests + smoke, GEMM after the shared-space fix, merged / unmerged / dedup bench, attention bwd EARLY_S, parity reports
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r3_tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/r3_tests.log
MOME_BUILD_CACHED=1 python __graft_entry__.py smoke > $O/r3_smoke.log 2>&1; echo "smoke rc=$?"; grep "smoke" $O/r3_smoke.log
python tools/gemm_bench.py > $O/r3_gb.log 2>&1; cat $O/r3_gb.log
MOME_GEMM_DEBUG=32 python tools/gemm_bench.py --only DGELU > $O/r3_gb_aux2.log 2>&1; cat $O/r3_gb_aux2.log
MOME_LN_BWD_VARIANT=0 python tools/row_bench.py > $O/r3_row_v0.log 2>&1; cat $O/r3_row_v0.log
MOME_LN_BWD_VARIANT=1 python tools/row_bench.py > $O/r3_row_v1.log 2>&1; cat $O/r3_row_v1.log
python tools/attn_bench.py --check --tc-bwd 2 --iters 20 > $O/r3_attn_bwd2.log 2>&1; echo "attn tc-bwd=2 rc=$?"; tail -12 $O/r3_attn_bwd2.log
python tools/attn_bench.py --tc-bwd 1 --iters 20 > $O/r3_attn_bwd1.log 2>&1; tail -8 $O/r3_attn_bwd1.log
python bench.py --steps 10 --warmup 3 > $O/r3_bench_merged.log 2>&1; tail -c 300 $O/r3_bench_merged.log
python bench.py --steps 10 --warmup 3 --no-merge --no-cpu-baseline --no-block-bench > $O/r3_bench_nomerge.log 2>&1; tail -c 300 $O/r3_bench_nomerge.log
python bench.py --steps 10 --warmup 3 --dedup --no-cpu-baseline --no-block-bench > $O/r3_bench_dedup.log 2>&1; tail -c 300 $O/r3_bench_dedup.log
python tools/parity_report.py --model vlmo_unit --batch 3 --out $O/parity_unit.json > $O/r3_parity_unit.log 2>&1; tail -3 $O/r3_parity_unit.log
python tools/parity_report.py --model vlmo_base --batch 2 --lengths full --out $O/parity_base.json > $O/r3_parity_base.log 2>&1; tail -3 $O/r3_parity_base.log
python tools/parity_report.py --model vlmo_large --batch 2 --lengths full --out $O/parity_large.json > $O/r3_parity_large.log 2>&1; tail -3 $O/r3_parity_large.log
python tools/parity_report.py --model vlmo_base --vqa480 --batch 2 --out $O/parity_vqa480.json > $O/r3_parity_vqa480.log 2>&1; tail -3 $O/r3_parity_vqa480.log
