#!/bin/bash
# r03r: phase timeline of the GI kernels after the shadow sub-ray move (r03q)
mkdir -p gpurun_out
timeout 300 python tools/gi_timeline.py > gpurun_out/r03r_gi_timeline.txt 2>&1
tail -45 gpurun_out/r03r_gi_timeline.txt
