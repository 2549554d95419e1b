#!/bin/bash
# r03v: gi_gen_trace0 at 6 resident CTAs per SM (40 registers) against 5 (48)
mkdir -p gpurun_out
for cfg in "VXPT_LIB=libvxpt.so" "VXPT_LIB=libvxpt_g6.so" "VXPT_LIB=libvxpt.so" "VXPT_LIB=libvxpt_g6.so"; do
  env $cfg timeout 120 python tools/gi_probe.py 20 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print(d['env'], {k:(round(d[k]['ms'],4), round(d[k]['frac_l2'],3)) for k in ('diffuse',)})"
done | tee gpurun_out/r03v_gi_probe.txt
