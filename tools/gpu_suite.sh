#!/bin/bash
# full GPU validation + bench lines (one gpurun call): tests, bench (bwe / noisybwe), results under gpurun_out/
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/suite_pytest.txt 2>&1
tail -6 gpurun_out/suite_pytest.txt | cut -c1-300
timeout 500 python bench.py --steps 10 --warmup 3 ${BENCH_FLAGS} > gpurun_out/suite_bench.json 2> gpurun_out/suite_bench.err
timeout 300 python bench.py --workload noisybwe --steps 10 --no-micro --no-cpu-baseline --no-eager-baseline > gpurun_out/suite_bench_noisy.json 2> gpurun_out/suite_bench_noisy.err
cut -c1-260 gpurun_out/suite_bench.json gpurun_out/suite_bench_noisy.json
