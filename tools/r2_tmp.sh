#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/p20_pytest.txt 2>&1
tail -5 gpurun_out/p20_pytest.txt | cut -c1-300
for v in 0 1; do echo "== VBX_D_PASS_STREAMS=$v" >> gpurun_out/p20_bench.txt; VBX_D_PASS_STREAMS=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-micro 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['ms_per_step'], d['value'], d['e2e']['losses'])" >> gpurun_out/p20_bench.txt; done
cat gpurun_out/p20_bench.txt
timeout 300 python tools/step_timeline.py graph > gpurun_out/p20_timeline_graph.txt 2>&1; rm -f gpurun_out/step_trace.json
head -14 gpurun_out/p20_timeline_graph.txt | tail -12
