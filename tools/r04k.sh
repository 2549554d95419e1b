#!/bin/bash
# r04k: y sweep of df_xy_dpx on the high bytes of the 16-bit lanes (three ALU-pipe instructions per row) against clean lanes (five,
# libvxpt_ycl.so = -DVXPT_DF_Y_CLEAN_LANES); then the whole GPU suite, smoke and the bench line on the final library
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_df_step_field.py -m gpu -x -q -k "df or distance or step" 2>&1 | tail -2 | tee gpurun_out/r04k_pytest_df.txt
for lib in libvxpt.so libvxpt_ycl.so libvxpt.so libvxpt_ycl.so; do
  for wld in plains city; do
    echo -n "$lib $wld "; VXPT_PROBE_WORLD=$wld VXPT_LIB=$lib timeout 120 python tools/df_probe.py 40 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print('algo1', d['algo1'])"
  done
done 2>&1 | tee gpurun_out/r04k_df_probe.txt
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r04k_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r04k_smoke.log
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r04k_bench_n1.json 2> gpurun_out/r04k_bench_n1.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r04k_bench_n1.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['df_build_ms'], d['roofline_all']['df_build']['frac'])"
