#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/p2_pytest.txt 2>&1
tail -5 gpurun_out/p2_pytest.txt
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/p2_bench.json 2> gpurun_out/p2_bench.err
timeout 200 python bench.py --workload noisybwe --steps 10 --no-micro --no-cpu-baseline --no-eager-baseline > gpurun_out/p2_bench_noisy.json 2> gpurun_out/p2_bench_noisy.err
cut -c1-400 gpurun_out/p2_bench.json gpurun_out/p2_bench_noisy.json
