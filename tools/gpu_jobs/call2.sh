#!/bin/bash
# GPU job: tests, parity reports, GEMM knob sweep, bench (all workloads), reference eager competitor, ncu of GELU/DGELU
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r2_tests1.log 2>&1; echo "tests rc=$?" 
python tools/parity_report.py --model vlmo_unit --batch 3 --out $O/parity_unit.json > $O/r2_parity_unit.log 2>&1; tail -3 $O/r2_parity_unit.log
python tools/parity_report.py --model vlmo_base --batch 2 --lengths full --out $O/parity_base.json > $O/r2_parity_base.log 2>&1; tail -3 $O/r2_parity_base.log
for dbg in 0 1 8 16 24; do
  MOME_GEMM_DEBUG=$dbg python tools/gemm_bench.py --only "qkv   fwd,GELU,proj  fwd" > $O/r2_gb_dbg$dbg.log 2>&1; cat $O/r2_gb_dbg$dbg.log
done
python tools/gemm_bench.py > $O/r2_gb_all.log 2>&1
python bench.py --steps 10 --warmup 3 > $O/r2_bench1.log 2>&1; tail -c 600 $O/r2_bench1.log
timeout 400 python tools/torch_eager_gpu.py --steps 3 --warmup 2 > $O/r2_ref_eager_pretrain.log 2>&1; tail -1 $O/r2_ref_eager_pretrain.log
python bench.py --workload vqa480 --steps 8 --warmup 3 > $O/r2_bench_vqa480.log 2>&1; tail -c 400 $O/r2_bench_vqa480.log
python bench.py --workload itc4096 --steps 8 --warmup 3 > $O/r2_bench_itc4096.log 2>&1; tail -c 400 $O/r2_bench_itc4096.log
python bench.py --model vlmo_large --steps 6 --warmup 3 > $O/r2_bench_large.log 2>&1; tail -c 400 $O/r2_bench_large.log
python tools/gemm_bench.py --only "dgrad DGELU   " --iters 3 > $O/r2_plain_dgelu.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_pair --launch-skip 3 --launch-count 1 -o $O/r2_dgelu python tools/gemm_bench.py --only "dgrad DGELU   " --iters 3 > $O/r2_ncu_dgelu.log 2>&1
python tools/gemm_bench.py --only "fwd  GELU" --iters 3 > $O/r2_plain_gelu.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_pair --launch-skip 3 --launch-count 1 -o $O/r2_gelu python tools/gemm_bench.py --only "fwd  GELU" --iters 3 > $O/r2_ncu_gelu.log 2>&1
ls -la $O/*.ncu-rep
