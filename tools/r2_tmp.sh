#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/step_timeline.py > gpurun_out/p17_timeline_eager.txt 2>&1
timeout 300 python tools/step_timeline.py graph > gpurun_out/p17_timeline_graph.txt 2>&1
rm -f gpurun_out/step_trace.json
cat gpurun_out/p17_timeline_graph.txt | tail -60
