#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1 2 4 8 16 31 30; do
  echo "== VBX_RU_DBG=$dbg" >> gpurun_out/p3_dbg.txt
  VBX_RU_DBG=$dbg timeout 120 python tools/ru_debug.py >> gpurun_out/p3_dbg.txt 2>&1
  echo "rc=$?" >> gpurun_out/p3_dbg.txt
done
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/ru_debug.py > gpurun_out/p3_sanitizer.txt 2>&1
grep -E "==|rc=|ok|err" gpurun_out/p3_dbg.txt | cut -c1-200
grep -E "Illegal|illegal|at 0x|PC|Error|ERROR SUMMARY" gpurun_out/p3_sanitizer.txt | head -20
# the rest of the GPU suite without the fused unit
VBX_FUSED_UNIT=0 timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_kernels.py::test_fused_residual_unit_matches_fp64_and_the_two_kernel_form --deselect tests/test_gpu_kernels.py::test_fused_residual_unit_through_autograd_matches_the_unfused_module > gpurun_out/p3_pytest.txt 2>&1
tail -15 gpurun_out/p3_pytest.txt
VBX_FUSED_UNIT=0 timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/p3_bench.json 2> gpurun_out/p3_bench.err
VBX_FUSED_UNIT=0 timeout 200 python bench.py --workload noisybwe --steps 10 --no-micro --no-cpu-baseline --no-eager-baseline > gpurun_out/p3_bench_noisy.json 2> gpurun_out/p3_bench_noisy.err
cut -c1-300 gpurun_out/p3_bench.json gpurun_out/p3_bench_noisy.json
