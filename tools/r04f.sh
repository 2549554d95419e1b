#!/bin/bash
# r04f: smallest share of the hit queue a gi_continue CTA takes (VXPT_GI_MIN_SHARE): 64 spreads small slabs over all SMs, 256 packs full chunks
mkdir -p gpurun_out
run() { timeout 300 python bench.py --gpus 1 --emulate $2 --steps 100 --warmup 10 --no-cpu-baseline --no-aux 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('min_share $1 emulate $2', round(d['value']), round(d['ms_per_step'],4), round(d['pass_ms']['diffuse'],4))"; }
for n in 2 4 8; do
  for ms in 64 128 256; do VXPT_GI_MIN_SHARE=$ms run $ms $n; done
done | tee gpurun_out/r04f_min_share.txt
