# r02h: GI continuation as separate kernels vs the CTA-staged kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "diffuse or gi or config4 or golden" > gpurun_out/r02h_pytest.log 2>&1; tail -4 gpurun_out/r02h_pytest.log
rm -f gpurun_out/r02h_gi_probe.jsonl
for env in "VXPT_GI_STAGED=1" "VXPT_GI_STAGED=0" "VXPT_GI_TRACE_CTAS=4" "VXPT_GI_TRACE_CTAS=8" "VXPT_GI_TRACE_CTAS=12"; do
  env $env timeout 120 python tools/gi_probe.py 20 >> gpurun_out/r02h_gi_probe.jsonl 2>> gpurun_out/r02h_gi_probe.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r02h_gi_probe.jsonl'):
    d=json.loads(l); print(d['env'], {k:(round(d[k]['ms'],4), round(d[k]['frac_l2'],3)) for k in ('primary','shadow','diffuse')})
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02h_launches_gi.csv python tools/gi_probe.py 3 > /dev/null 2>&1
grep -E "gi_" gpurun_out/r02h_launches_gi.csv | tail -6 | awk -F'","' '{print $5, $NF}' | cut -c1-200
