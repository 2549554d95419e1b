#!/bin/bash
# r03q: gi_continue with the first hit's sun-shadow sub-ray traced in stage D (late) against stage B (early); GI parity on the new order
mkdir -p gpurun_out
for lib in libvxpt.so libvxpt_early.so libvxpt.so libvxpt_early.so; do
  VXPT_LIB=$lib timeout 120 python tools/gi_probe.py 20 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print(d['env'], {k:(round(d[k]['ms'],4), round(d[k]['frac_l2'],3)) for k in ('diffuse',)})"
done | tee gpurun_out/r03q_gi_probe.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r03q_pytest.txt
