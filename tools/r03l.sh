# r03l: GI row sub-slabs again, now that both kernels fit an SM together (gen held at 4 CTAs per SM, one gi_continue CTA beside them)
mkdir -p gpurun_out
for env in "VXPT_GI_SLABS=1" "VXPT_GI_SLABS=2" "VXPT_GI_SLABS=4" "VXPT_GI_SLABS=8" "VXPT_GI_SLABS=4 VXPT_GI_SLAB_CTAS=2" "VXPT_GI_SLABS=8 VXPT_GI_SLAB_CTAS=2" "VXPT_GI_SLABS=4 VXPT_GI_GEN_PAD=0" "VXPT_GI_SLABS=6 VXPT_GI_SLAB_CTAS=1"; do
  env $env timeout 120 python tools/gi_probe.py 20 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print(d['env'], {k:(round(d[k]['ms'],4), round(d[k]['frac_l2'],3)) for k in ('diffuse',)})"
done
