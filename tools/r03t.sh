#!/bin/bash
# r03t: gi_continue (both sub-rays in stage D, results folded into the ray arrays) at 4 / 5 / 6 resident CTAs per SM; GI parity on the 5- and 6-CTA builds
mkdir -p gpurun_out
for cfg in "VXPT_LIB=libvxpt_c4.so" "VXPT_LIB=libvxpt_c5.so" "VXPT_LIB=libvxpt_c6.so" "VXPT_LIB=libvxpt_c4.so" "VXPT_LIB=libvxpt_c5.so" "VXPT_LIB=libvxpt_c6.so" "VXPT_LIB=libvxpt_c6.so VXPT_GI_CTAS=5"; do
  env $cfg timeout 120 python tools/gi_probe.py 20 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print(d['env'], {k:(round(d[k]['ms'],4), round(d[k]['frac_l2'],3)) for k in ('diffuse',)})"
done | tee gpurun_out/r03t_gi_probe.txt
cp voxelpathtracer_b200/libvxpt.so /tmp/libvxpt_keep.so
for m in 5 6; do
  cp voxelpathtracer_b200/libvxpt_c$m.so voxelpathtracer_b200/libvxpt.so
  echo "pytest with MINB=$m" | tee -a gpurun_out/r03t_pytest.txt
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "diffuse or gi or frame or golden" 2>&1 | tail -3 | tee -a gpurun_out/r03t_pytest.txt
done
cp /tmp/libvxpt_keep.so voxelpathtracer_b200/libvxpt.so
