#!/bin/bash
# r03u: gi_continue with even shares of the hit queue per CTA (no cursor), 5 CTAs per SM; whole frame and one eighth of it (an 8-GPU slab)
mkdir -p gpurun_out
for args in "20" "20 1920 136" "20"; do
  timeout 120 python tools/gi_probe.py $args 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print(d['resolution'], {k:(round(d[k]['ms'],4), round(d[k]['frac_l2'],3)) for k in ('primary','shadow','diffuse')})"
done | tee gpurun_out/r03u_gi_probe.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r03u_pytest.txt
