# r02i: streaming z sweep + half-word y sweep (distance field); GI bounce chunk size
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_df_step_field.py tests/test_per_frame_edits.py -m gpu -x -q -k "df or distance or step or edit or diffuse or config4" > gpurun_out/r02i_pytest.log 2>&1; tail -4 gpurun_out/r02i_pytest.log
for env in "VXPT_DF_Z=1 VXPT_DF_YHALF=1" "VXPT_DF_Z=0 VXPT_DF_YHALF=0" "VXPT_DF_Z=1 VXPT_DF_YHALF=0" "VXPT_DF_Z=0 VXPT_DF_YHALF=1" "VXPT_DF_XY_CTAS=148" "VXPT_DF_XY_CTAS=384"; do
  echo "$env"; env $env timeout 120 python tools/df_probe.py 30 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print('  algo1', d['algo1'])"
done
for env in "VXPT_GI_STAGED=1" "VXPT_GI_BOUNCE_NR=1" "VXPT_GI_BOUNCE_NR=2" "VXPT_GI_BOUNCE_NR=1 VXPT_GI_TRACE_CTAS=4"; do
  env $env timeout 120 python tools/gi_probe.py 20 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print(d['env'], {k:(round(d[k]['ms'],4), round(d[k]['frac_l2'],3)) for k in ('primary','shadow','diffuse')})"
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02i_launches_df.csv python tools/df_probe.py 3 > /dev/null 2>&1
grep -E "df_|pack" gpurun_out/r02i_launches_df.csv | head -4 | awk -F'","' '{print $5, $NF}' | cut -c1-200
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02i_launches_gi.csv python tools/gi_probe.py 3 > /dev/null 2>&1
grep -E "gi_" gpurun_out/r02i_launches_gi.csv | tail -6 | awk -F'","' '{print $5, $NF}' | cut -c1-200
for k in df_xy_dpx df_z_stream; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_r02i_$k python tools/df_probe.py 3 > gpurun_out/r02i_ncu_$k.log 2>&1
done
