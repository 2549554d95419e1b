#!/bin/bash
# r03x: primary / shadow kernels at 6 resident CTAs per SM (40 registers; the shadow kernel has 44 uncapped); reflection parity on the capped kernels
mkdir -p gpurun_out
for cfg in "VXPT_LIB=libvxpt.so" "VXPT_LIB=libvxpt_t6.so" "VXPT_LIB=libvxpt.so" "VXPT_LIB=libvxpt_t6.so"; do
  env $cfg timeout 120 python tools/gi_probe.py 20 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print(d['env'], {k:(round(d[k]['ms'],4), round(d[k]['frac_l2'],3)) for k in ('primary','shadow','diffuse')})"
done | tee gpurun_out/r03x_probe.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_z_material_extras.py -m gpu -x -q -k "reflection" 2>&1 | tail -3 | tee gpurun_out/r03x_pytest.txt
