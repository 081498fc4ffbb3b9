#!/bin/bash
# Round-2 measurement batch (one gpurun call, one GPU): everything profiles/README.md cites except the three
# `ncu --set full` captures (tools/r2_ncu.sh) and the multi-GPU lines (tools/r2_multi.sh).
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > $O/r2_pytest_gpu.txt 2>&1; tail -3 $O/r2_pytest_gpu.txt | cut -c1-200
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2_bench_1gpu.json 2> $O/r2_bench_1gpu.err
timeout 300 python bench.py --workload noisybwe --steps 20 --warmup 5 --no-micro > $O/r2_bench_noisybwe_1gpu.json 2> $O/r2_bench_noisybwe_1gpu.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference_arm.json 2> /dev/null
cut -c1-200 $O/r2_bench_1gpu.json $O/r2_bench_noisybwe_1gpu.json $O/r2_bench_reference_arm.json
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches.csv python bench.py --steps 1 --profile > /dev/null 2>&1
python tools/summarize_launches.py $O/r2_launches.csv > $O/r2_launch_list_summary.txt 2>&1
timeout 300 python tools/layer_table.py > $O/r2_layer_roofline_table.md 2> /dev/null
timeout 300 python tools/step_timeline.py graph > $O/r2_step_timeline.txt 2>&1; rm -f $O/step_trace.json
timeout 400 python bench.py --sweep --sweep-budget-s 200 > $O/r2_sweep.json 2> $O/r2_sweep.err; mv $O/conv_sweep.md $O/r2_conv_sweep_vs_cudnn.md 2>/dev/null
cut -c1-400 $O/r2_sweep.json
ls -la $O | grep r2_ | head -40
