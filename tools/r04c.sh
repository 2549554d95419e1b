#!/bin/bash
# r04c: GI pass on slabs of 1/1, 1/2, 1/4, 1/8 of the 1080p frame: library of r03o (before the continuation rework) against the final one
mkdir -p gpurun_out
for rows in 1080 544 272 136; do
  for lib in libvxpt_old.so libvxpt.so; do
    VXPT_LIB=$lib timeout 120 python tools/gi_probe.py 20 1920 $rows 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print('$lib', d['resolution'], {k:round(d[k]['ms'],4) for k in ('primary','shadow','diffuse')})"
  done
done | tee gpurun_out/r04c_gi_slabs.txt
