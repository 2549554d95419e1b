# r02d: new distance-field kernels (persistent xy sweep, fused step field, PDL) and GI sub-slab pipelining
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest.log 2>&1; tail -6 gpurun_out/r02d_pytest.log
timeout 120 python tools/df_probe.py 30 > gpurun_out/r02d_df_probe.json 2> gpurun_out/r02d_df_probe.err; cat gpurun_out/r02d_df_probe.json
for c in 148 192 384; do VXPT_DF_XY_CTAS=$c timeout 120 python tools/df_probe.py 30 2>&1 | cut -c1-400; done
for sl in 1 2 3 4; do VXPT_GI_SLABS=$sl timeout 120 python tools/gi_probe.py 20 >> gpurun_out/r02d_gi_probe.jsonl 2>> gpurun_out/r02d_gi_probe.err; done
python - <<'PY'
import json
for l in open('gpurun_out/r02d_gi_probe.jsonl'):
    d=json.loads(l); print(d['env'], {k:(round(d[k]['ms'],4), round(d[k]['frac_l2'],3)) for k in ('primary','shadow','diffuse')})
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02d_launches_df.csv python tools/df_probe.py 3 > /dev/null 2>&1
grep -E "df_|pack" gpurun_out/r02d_launches_df.csv | tail -8 | cut -c1-260
for k in df_xy_dpx df_z_dpx; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_r02d_$k python tools/df_probe.py 3 > gpurun_out/r02d_ncu_$k.log 2>&1
done
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r02d_bench_n1.json 2> gpurun_out/r02d_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r02d_bench_n1.json')); print(d['value'], d['ms_per_step'], d['pass_ms'], d['roofline']['frac'], d['roofline_all']['df_build'], d['e2e']['value'])"; tail -3 gpurun_out/r02d_bench_n1.err
