#!/bin/bash
# r04h: (a) frame pipes on one GPU (1 / 2 / 3 / 4), (b) primary / shadow kernels with 128-thread CTAs (32x4 pixels) against 256 (32x8),
# (c) one rank's share of a 4- / 8-way sharded frame under the driver's 20-step flags with a pipe count that divides the step count
#     (4, 5, 10) against the default 8
mkdir -p gpurun_out
cp voxelpathtracer_b200/libvxpt.so /tmp/final.so
run() { timeout 300 python bench.py --gpus 1 $2 --no-cpu-baseline --no-aux 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); p=d['pass_ms']; print('$1 | $2 |', round(d['value']), round(d['ms_per_step'],4), round(p['primary'],4), round(p['shadow'],4), round(p['diffuse'],4))"; }
{
for p in 2 1 3 4; do run cta256 "--steps 100 --warmup 10 --pipes $p"; done
cp voxelpathtracer_b200/libvxpt_cta128.so voxelpathtracer_b200/libvxpt.so
run cta128 "--steps 100 --warmup 10 --pipes 2"
run cta128 "--steps 100 --warmup 10 --pipes 3"
cp /tmp/final.so voxelpathtracer_b200/libvxpt.so
for rep in 1 2; do for p in 8 10 5 4; do run cta256 "--emulate 8 --steps 20 --warmup 5 --pipes $p"; done; done
for p in 8 10 5 4; do run cta256 "--emulate 4 --steps 20 --warmup 5 --pipes $p"; done
cp voxelpathtracer_b200/libvxpt_cta128.so voxelpathtracer_b200/libvxpt.so
for p in 8 10; do run cta128 "--emulate 8 --steps 20 --warmup 5 --pipes $p"; done
cp /tmp/final.so voxelpathtracer_b200/libvxpt.so
} 2>&1 | tee gpurun_out/r04h_pipes_cta.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "primary or shadow" 2>&1 | tail -2 | tee gpurun_out/r04h_pytest.txt
