#!/bin/bash
# r03z: denoiser kernels with register caps: 4 CTAs per SM (64 registers) and 5 (48) against uncapped (62-80)
mkdir -p gpurun_out
for lib in libvxpt.so libvxpt_d4.so libvxpt_d5.so libvxpt.so libvxpt_d4.so libvxpt_d5.so; do
  VXPT_LIB=$lib timeout 120 python tools/denoise_probe.py 10 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$lib', {k: round(v['ms'],4) for k,v in d['passes'].items() if k not in ('material','reflection')})"
done | tee gpurun_out/r03z_denoise.txt
