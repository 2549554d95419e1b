#!/bin/bash
# r03s: gi_continue at 5 resident CTAs per SM (48 registers): 740 chunk slots against the bench frame's 714 chunks
mkdir -p gpurun_out
for cfg in "VXPT_LIB=libvxpt.so" "VXPT_LIB=libvxpt_c5.so VXPT_GI_CTAS=5" "VXPT_LIB=libvxpt_c5.so VXPT_GI_CTAS=4" "VXPT_LIB=libvxpt.so" "VXPT_LIB=libvxpt_c5.so VXPT_GI_CTAS=5"; do
  env $cfg timeout 120 python tools/gi_probe.py 20 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print(d['env'], {k:(round(d[k]['ms'],4), round(d[k]['frac_l2'],3)) for k in ('diffuse',)})"
done | tee gpurun_out/r03s_gi_probe.txt
