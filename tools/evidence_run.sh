#!/bin/bash
# Evidence of one build: ncu launch list of a 1048-pair slice of the corpus step (4 sub-batches), then ncu --set full of the
# first-scale kernels of each metric.   usage: tools/evidence_run.sh <tag>;  read here with tools/launch_table.py / ncu_summary.py
tag=${1:-ev}
mkdir -p gpurun_out
python bench.py --groups 131 --ncu > gpurun_out/${tag}_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches_raw.csv python bench.py --groups 131 --ncu > gpurun_out/${tag}_ncu.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/${tag}_ncu_plain.log
bash tools/ncu_run.sh ${tag}
