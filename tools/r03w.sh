#!/bin/bash
# r03w: re-queued reflection pass with register caps: refl_gen_trace at 5 CTAs per SM (48 registers), refl_shade at 4 (64)
mkdir -p gpurun_out
for lib in libvxpt.so libvxpt_r5.so libvxpt_r5s4.so libvxpt.so libvxpt_r5.so libvxpt_r5s4.so; do
  VXPT_LIB=$lib timeout 120 python tools/denoise_probe.py 10 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$lib', round(d['passes']['reflection']['ms'],4))"
done | tee gpurun_out/r03w_reflection.txt
