#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 2 3 4 5 6 7 8; do timeout 60 ./tools/tma_probe $v >> gpurun_out/p4_tma.txt 2>&1; echo "rc=$?" >> gpurun_out/p4_tma.txt; done
cat gpurun_out/p4_tma.txt | cut -c1-220
VBX_FUSED_UNIT=0 timeout 600 python -m pytest tests -m gpu -q -k "pqmf or trajectory" > gpurun_out/p4_pytest.txt 2>&1
tail -5 gpurun_out/p4_pytest.txt; cat gpurun_out/trajectory_tc.txt gpurun_out/trajectory_fma.txt
