#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/ru_debug.py 1 32 256 1 > gpurun_out/p5_dbg.txt 2>&1; echo "rc=$?" >> gpurun_out/p5_dbg.txt
timeout 120 python tools/ru_debug.py 3 64 1000 9 >> gpurun_out/p5_dbg.txt 2>&1; echo "rc=$?" >> gpurun_out/p5_dbg.txt
grep -E "ok|err|rc=|rror" gpurun_out/p5_dbg.txt | cut -c1-200
timeout 900 python -m pytest tests -m gpu -q -x -k "fused_residual" > gpurun_out/p5_pytest_ru.txt 2>&1
tail -12 gpurun_out/p5_pytest_ru.txt | cut -c1-300
if grep -q "passed" gpurun_out/p5_pytest_ru.txt && ! grep -q "failed" gpurun_out/p5_pytest_ru.txt; then
  timeout 900 python -m pytest tests -m gpu -q > gpurun_out/p5_pytest.txt 2>&1
  tail -8 gpurun_out/p5_pytest.txt | cut -c1-300
  timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/p5_bench.json 2> gpurun_out/p5_bench.err
  cut -c1-300 gpurun_out/p5_bench.json
fi
cat gpurun_out/trajectory_tc.txt gpurun_out/trajectory_fma.txt 2>/dev/null
