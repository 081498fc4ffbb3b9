#!/bin/bash
# scratch batch: chain-fusion checks, bench, sweep
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "gate_stage" > $O/c1_pytest_gate.txt 2>&1; tail -3 $O/c1_pytest_gate.txt | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "fused_chain or training_step_matches" > $O/c1_pytest_chain.txt 2>&1; tail -3 $O/c1_pytest_chain.txt | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-eager-baseline --no-cpu-baseline > $O/c1_bench.json 2> $O/c1_bench.err; cut -c1-260 $O/c1_bench.json
VBX_CHAIN_FUSION=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-eager-baseline --no-cpu-baseline --no-micro > $O/c1_bench_unfused.json 2> $O/c1_bench_unfused.err; cut -c1-260 $O/c1_bench_unfused.json
timeout 1200 python -m pytest tests -m gpu -q > $O/c1_pytest_gpu.txt 2>&1; tail -3 $O/c1_pytest_gpu.txt | cut -c1-300
timeout 500 python bench.py --sweep --sweep-budget-s 300 > $O/c1_sweep.json 2> $O/c1_sweep.err; mv $O/conv_sweep.md $O/r2_conv_sweep_vs_cudnn.md 2>/dev/null
cut -c1-400 $O/c1_sweep.json; tail -3 $O/c1_sweep.err
