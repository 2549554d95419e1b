# r02s: distance field: one CTA per slice (4 per SM), batched y sweep, parallel carries in the z sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_df_step_field.py tests/test_per_frame_edits.py -m gpu -x -q -k "df or distance or step or edit" > gpurun_out/r02s_pytest.log 2>&1; tail -3 gpurun_out/r02s_pytest.log
python tools/df_timeline.py
for env in "VXPT_DF_XY_CTAS=384" "VXPT_DF_XY_CTAS=296" "VXPT_PROBE_WORLD=city"; do
  echo "$env"; env $env timeout 120 python tools/df_probe.py 30 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print('  algo1', d['algo1'])"
done
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02s_launches_df.csv python tools/df_probe.py 3 > /dev/null 2>&1
grep -E "df_xy|df_z_dpx" gpurun_out/r02s_launches_df.csv | head -4 | awk -F'","' '{print substr($5,1,30), $(NF-2), $NF}' | cut -c1-200
