#!/bin/bash
mkdir -p gpurun_out
VBX_WS_TS=0 python tools/wslab_probe.py > gpurun_out/p25.txt 2>&1
VBX_WS_TS=1 python tools/wslab_probe.py >> gpurun_out/p25.txt 2>&1
cat gpurun_out/p25.txt
VBX_WS_TS=0 ncu --set full --clock-control none --import-source on -k regex:tc_wslab_kernel -s 3 -c 1 -o gpurun_out/p25_wslab_ss python tools/wslab_probe.py > /dev/null 2>&1
VBX_WS_TS=1 ncu --set full --clock-control none --import-source on -k regex:tc_wslab_kernel -s 3 -c 1 -o gpurun_out/p25_wslab_ts python tools/wslab_probe.py > /dev/null 2>&1
ls -la gpurun_out/p25*
