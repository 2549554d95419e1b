# r03m: GI sub-slabs, gen held at 3 CTAs per SM, last continuation with full CTAs
mkdir -p gpurun_out
for env in "VXPT_GI_SLABS=4" "VXPT_GI_SLABS=4 VXPT_GI_GEN_PAD=22000" "VXPT_GI_SLABS=4 VXPT_GI_GEN_PAD=22000 VXPT_GI_SLAB_CTAS=2" "VXPT_GI_SLABS=8 VXPT_GI_GEN_PAD=22000 VXPT_GI_SLAB_CTAS=1" "VXPT_GI_SLABS=3 VXPT_GI_GEN_PAD=22000 VXPT_GI_SLAB_CTAS=2"; do
  env $env timeout 120 python tools/gi_probe.py 20 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print(d['env'], {k:(round(d[k]['ms'],4), round(d[k]['frac_l2'],3)) for k in ('diffuse',)})"
done
