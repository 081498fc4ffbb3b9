#!/bin/bash
# scratch batch: where should the gate stage run
mkdir -p gpurun_out
O=gpurun_out
for ge in 13 15 0; do
VBX_GATE_EPILOGUE=$ge timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "gate_stage" 2>&1 | tail -1 | cut -c1-200
done
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "fused_chain" 2>&1 | tail -1 | cut -c1-300
B="python bench.py --steps 20 --warmup 5 --no-eager-baseline --no-cpu-baseline --no-micro"
for ge in 13 15 0 9 5 1; do
echo "VBX_GATE_EPILOGUE=$ge"; VBX_GATE_EPILOGUE=$ge timeout 300 $B 2> /dev/null | grep -o '"ms_per_step": [0-9.]*'
done
echo "VBX_CHAIN_FUSION=0"; VBX_CHAIN_FUSION=0 timeout 300 $B 2> /dev/null | grep -o '"ms_per_step": [0-9.]*'
echo "default again"; timeout 300 $B 2> /dev/null | grep -o '"ms_per_step": [0-9.]*'
