#!/bin/bash
# round-2 first GPU call: baselines before any kernel work
mkdir -p gpurun_out
nproc > gpurun_out/p1_host.txt; nvidia-smi -L >> gpurun_out/p1_host.txt; lscpu | head -20 >> gpurun_out/p1_host.txt
python tools/narrow_check.py > gpurun_out/p1_narrow_default.txt 2>&1
VBX_TC_PS_VEC=1 python tools/narrow_check.py > gpurun_out/p1_narrow_vec.txt 2>&1
python tools/gpu_eager_probe.py 32 --cpu > gpurun_out/p1_eager.txt 2>&1
tail -3 gpurun_out/p1_narrow_default.txt gpurun_out/p1_narrow_vec.txt gpurun_out/p1_eager.txt
