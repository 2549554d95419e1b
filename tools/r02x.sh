# r02x: gi_continue stage B traces two groups per warp interleaved (two rays per lane)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_per_frame_edits.py -m gpu -x -q -k "diffuse or gi or config4 or golden or edit" > gpurun_out/r02x_pytest.log 2>&1; tail -3 gpurun_out/r02x_pytest.log
for lib in libvxpt.so libvxpt_c3.so; do
  VXPT_LIB=$lib timeout 120 python tools/gi_probe.py 20 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print('$lib', {k:(round(d[k]['ms'],4), round(d[k]['frac_l2'],3)) for k in ('primary','shadow','diffuse')})"
done
