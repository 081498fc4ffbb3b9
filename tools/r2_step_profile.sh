#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_suite.sh
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --profile > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/launches_r2.csv > gpurun_out/launch_summary_r2.txt 2>&1
head -45 gpurun_out/launch_summary_r2.txt
timeout 300 python tools/layer_table.py > gpurun_out/layer_table_r2.md 2> gpurun_out/layer_table_r2.err
head -3 gpurun_out/layer_table_r2.md
