#!/bin/bash
# r04d: does the reworked GI continuation (740 CTAs that hold 61 k of an SM's 64 k registers) crowd out the other frames' kernels in the
# pipelined bench?  Old library (r03o) against the final one, and the final one with 3 / 4 continuation CTAs per SM, 100-step runs, one box
mkdir -p gpurun_out
cp voxelpathtracer_b200/libvxpt.so /tmp/final.so
run() { timeout 300 python bench.py --gpus 1 --steps 100 --warmup 10 --no-cpu-baseline --no-aux 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', round(d['value']), round(d['ms_per_step'],4), d['pass_ms']['diffuse'])"; }
for rep in 1 2; do
  cp voxelpathtracer_b200/libvxpt_old.so voxelpathtracer_b200/libvxpt.so; run old
  cp /tmp/final.so voxelpathtracer_b200/libvxpt.so; run final
  VXPT_GI_CTAS=4 run final_ctas4
  VXPT_GI_CTAS=3 run final_ctas3
  VXPT_GI_CTAS=2 run final_ctas2
done | tee gpurun_out/r04d_pipelined.txt
cp /tmp/final.so voxelpathtracer_b200/libvxpt.so
