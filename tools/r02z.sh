# r02z: primary / shadow kernels capped at 32 registers (8 CTAs per SM)
mkdir -p gpurun_out
for lib in libvxpt.so libvxpt_t8.so; do
  VXPT_LIB=$lib timeout 120 python tools/gi_probe.py 20 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print('$lib', {k:(round(d[k]['ms'],4), round(d[k]['frac_l2'],3)) for k in ('primary','shadow','diffuse')})"
done
