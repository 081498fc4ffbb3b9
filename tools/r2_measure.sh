#!/bin/bash
# Round-2 measurement batch (one gpurun call, one GPU): everything profiles/README.md cites.
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > $O/r2_pytest_gpu.txt 2>&1; tail -3 $O/r2_pytest_gpu.txt | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2_bench_1gpu.json 2> $O/r2_bench_1gpu.err
timeout 300 python bench.py --workload noisybwe --steps 20 --warmup 5 --no-micro > $O/r2_bench_noisybwe_1gpu.json 2> $O/r2_bench_noisybwe_1gpu.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference_arm.json 2> /dev/null
cut -c1-200 $O/r2_bench_1gpu.json $O/r2_bench_noisybwe_1gpu.json $O/r2_bench_reference_arm.json
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches.csv python bench.py --steps 1 --profile > /dev/null 2>&1
python tools/summarize_launches.py $O/r2_launches.csv > $O/r2_launch_list_summary.txt 2>&1
timeout 300 python tools/layer_table.py > $O/r2_layer_roofline_table.md 2> /dev/null
ncu --set full --clock-control none --import-source on -k regex:ru_fwd_kernel -s 3 -c 1 -o $O/r2_ncu_ru_fwd_c32 python tools/ru_bench.py 32 11968 3 6 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:ru_fwd_kernel -s 3 -c 1 -o $O/r2_ncu_ru_fwd_c64 python tools/ru_bench.py 64 5984 9 6 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_slab_kernel -c 1 -o $O/r2_ncu_slab_melgan4 python -c "
import sys; sys.path.insert(0,'.')
import torch
from vibravox_b200 import ops
g = ops.ConvGeom(1024, 1024, 41, 4, 1, 20, 0, 4)
x = torch.randn(32, 1024, 748, device='cuda'); w = torch.randn(1024, 256, 41, device='cuda') * 0.01; b = torch.zeros(1024, device='cuda')
for _ in range(3): ops.conv_fwd(x, w, g, bias=b, slope=0.2)
torch.cuda.synchronize()" > /dev/null 2>&1
python tools/ru_bench.py > $O/r2_ru_bench.txt 2>&1; RU_TRAIN=1 python tools/ru_bench.py >> $O/r2_ru_bench.txt 2>&1
python tools/wg_bench.py > $O/r2_wg_bench.txt 2>&1
python tools/ru_timeline.py 32 11968 3 > $O/r2_ru_timeline_c32.txt 2>&1
python tools/ru_timeline.py 64 5984 9 > $O/r2_ru_timeline_c64.txt 2>&1
timeout 300 python tools/step_timeline.py graph > $O/r2_step_timeline.txt 2>&1; rm -f $O/step_trace.json
timeout 900 python bench.py --sweep > $O/r2_sweep.json 2> $O/r2_sweep.err; mv $O/conv_sweep.md $O/r2_conv_sweep_vs_cudnn.md 2>/dev/null
cat $O/r2_sweep.json | cut -c1-400
ls -la $O | grep r2_
