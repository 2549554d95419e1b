#!/bin/bash
# r04g: gi_continue deals a short queue to ~355 CTAs: emulated 1 / 2 / 4 / 8-way shares against the r03o library, GI parity
mkdir -p gpurun_out
cp voxelpathtracer_b200/libvxpt.so /tmp/final.so
run() { timeout 300 python bench.py --gpus 1 $2 --steps 100 --warmup 10 --no-cpu-baseline --no-aux 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1 $2', round(d['value']), round(d['ms_per_step'],4), round(d['pass_ms']['diffuse'],4))"; }
for e in "" "--emulate 2" "--emulate 4" "--emulate 8"; do
  cp voxelpathtracer_b200/libvxpt_old.so voxelpathtracer_b200/libvxpt.so; run old "$e"
  cp /tmp/final.so voxelpathtracer_b200/libvxpt.so; run final "$e"
done | tee gpurun_out/r04g_auto_share.txt
cp /tmp/final.so voxelpathtracer_b200/libvxpt.so
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/r04g_pytest.txt
