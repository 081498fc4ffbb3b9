#!/bin/bash
# multi-GPU bench lines (one box, N ranks over NCCL): usage tools/r2_multi.sh N
N=$1
mkdir -p gpurun_out
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $T bench.py --gpus $N --steps 20 --warmup 5 --no-micro > $O/r2_bench_${N}gpu.json 2> $O/r2_bench_${N}gpu.err
cut -c1-300 $O/r2_bench_${N}gpu.json; tail -2 $O/r2_bench_${N}gpu.err | cut -c1-300
timeout 600 $T bench.py --gpus $N --workload noisybwe --steps 20 --warmup 5 --no-micro > $O/r2_bench_noisybwe_${N}gpu.json 2> $O/r2_bench_noisybwe_${N}gpu.err
cut -c1-300 $O/r2_bench_noisybwe_${N}gpu.json; tail -2 $O/r2_bench_noisybwe_${N}gpu.err | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-micro --no-cpu-baseline --no-eager-baseline > $O/r2_bench_1gpu_samebox_${N}.json 2> /dev/null
cut -c1-300 $O/r2_bench_1gpu_samebox_${N}.json
