#!/bin/bash
# ncu launch list of one eager VQA-480 step
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
python bench.py --workload vqa480 --ncu-step > $O/r24_plain.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r24_launches_vqa480.csv python bench.py --workload vqa480 --ncu-step > $O/r24_ncu.log 2>&1; echo "launch list rc=$?"
