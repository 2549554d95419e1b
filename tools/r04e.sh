#!/bin/bash
# r04e: rank 0's share of a 2- / 4- / 8-way sharded frame on one GPU (bench --emulate), old library (r03o) against the final one
mkdir -p gpurun_out
cp voxelpathtracer_b200/libvxpt.so /tmp/final.so
run() { timeout 300 python bench.py --gpus 1 --emulate $2 --steps 100 --warmup 10 --no-cpu-baseline --no-aux 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1 emulate $2', round(d['value']), round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['pass_ms'].items() if isinstance(v,float)})"; }
for n in 2 4 8; do
  for rep in 1 2; do
    cp voxelpathtracer_b200/libvxpt_old.so voxelpathtracer_b200/libvxpt.so; run old $n
    cp /tmp/final.so voxelpathtracer_b200/libvxpt.so; run final $n
  done
done | tee gpurun_out/r04e_emulate.txt
cp /tmp/final.so voxelpathtracer_b200/libvxpt.so
