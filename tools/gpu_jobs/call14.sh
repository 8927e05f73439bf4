#!/bin/bash
# GPU job 14 (r6): final single-GPU collection: tests, smoke, micro-benchmarks, every bench workload, launch list, ncu of the new attention kernels, parity reports
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
nvidia-smi -L > $O/r6_gpu.txt
python -m pytest tests -m gpu -q > $O/r6_tests.log 2>&1; echo "tests rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" $O/r6_tests.log | tail -8
MOME_BUILD_CACHED=1 python __graft_entry__.py smoke > $O/r6_smoke.log 2>&1; echo "smoke rc=$?"; grep "smoke" $O/r6_smoke.log
python tools/gemm_bench.py > $O/r6_gb.log 2>&1; echo "gemm rc=$?"
python tools/row_bench.py > $O/r6_row.log 2>&1; echo "row rc=$?"
python tools/attn_bench.py --check --tc-bwd p --iters 20 > $O/r6_attn.log 2>&1; echo "attn rc=$?"; tail -9 $O/r6_attn.log | cut -c1-110
python tools/attn_bench.py --check --tc-bwd p --iters 10 --long > $O/r6_attn_long.log 2>&1; echo "attn long rc=$?"; tail -5 $O/r6_attn_long.log | cut -c1-110
python tools/attn_bench.py --tc-bwd 1 --iters 20 --only fused > $O/r6_attn_first_kernel.log 2>&1
python bench.py --steps 10 --warmup 3 > $O/r6_bench.log 2>&1; echo "bench rc=$?"; tail -c 400 $O/r6_bench.log
python bench.py --steps 10 --warmup 3 --no-merge --no-cpu-baseline --no-block-bench > $O/r6_bench_nomerge.log 2>&1; echo "nomerge rc=$?"
python bench.py --steps 10 --warmup 3 --dedup --no-cpu-baseline --no-block-bench > $O/r6_bench_dedup.log 2>&1; echo "dedup rc=$?"
python bench.py --workload vqa480 --steps 8 --warmup 3 > $O/r6_bench_vqa480.log 2>&1; echo "vqa rc=$?"; tail -c 300 $O/r6_bench_vqa480.log
python bench.py --workload itc4096 --steps 8 --warmup 3 > $O/r6_bench_itc4096.log 2>&1; echo "itc rc=$?"; tail -c 300 $O/r6_bench_itc4096.log
python bench.py --model vlmo_large --steps 6 --warmup 3 > $O/r6_bench_large.log 2>&1; echo "large rc=$?"; tail -c 300 $O/r6_bench_large.log
python bench.py --impl reference --steps 2 --warmup 1 > $O/r6_bench_reference.log 2>&1; echo "reference rc=$?"; tail -c 300 $O/r6_bench_reference.log
python tools/parity_report.py --model vlmo_base --batch 2 --lengths full --out $O/parity_base.json > $O/r6_parity_base.log 2>&1; tail -2 $O/r6_parity_base.log
python tools/parity_report.py --model vlmo_base --vqa480 --batch 2 --out $O/parity_vqa480.json > $O/r6_parity_vqa480.log 2>&1; tail -2 $O/r6_parity_vqa480.log
python bench.py --ncu-step > $O/r6_plain.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r6_launches.csv python bench.py --ncu-step > $O/r6_ncu.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc_pipe_kernel -s 2 -c 1 -f -o $O/r6_attn_bwd_pipe python tools/attn_bench.py --only fused --tc-bwd p --iters 1 > $O/r6_ncu_bwd.log 2>&1; echo "ncu bwd rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc_long_kernel -s 2 -c 1 -f -o $O/r6_attn_fwd_long python tools/attn_bench.py --long --only vqa --tc-bwd p --iters 1 > $O/r6_ncu_fwd_long.log 2>&1; echo "ncu fwd long rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc_pipe_kernel -s 2 -c 1 -f -o $O/r6_attn_bwd_pipe_long python tools/attn_bench.py --long --only vqa --tc-bwd p --iters 1 > $O/r6_ncu_bwd_long.log 2>&1; echo "ncu bwd long rc=$?"
