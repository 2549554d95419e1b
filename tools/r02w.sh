# r02w: gi_continue capped at 64 registers (4 CTAs per SM: every chunk of the slab resident at once), gi_gen_trace0 at 51
mkdir -p gpurun_out
for lib in libvxpt.so libvxpt_c4.so libvxpt_c4g5.so; do
  for ctas in 4 5; do
  VXPT_LIB=$lib VXPT_GI_CTAS=$ctas timeout 120 python tools/gi_probe.py 20 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print('$lib', d['env'], {k:(round(d[k]['ms'],4), round(d[k]['frac_l2'],3)) for k in ('primary','shadow','diffuse')})"
  done
done
