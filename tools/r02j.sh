# r02j: batched column sweeps in the distance-field kernels; compute-sanitizer
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_df_step_field.py tests/test_per_frame_edits.py -m gpu -x -q -k "df or distance or step or edit or diffuse" > gpurun_out/r02j_pytest.log 2>&1; tail -3 gpurun_out/r02j_pytest.log
for env in "VXPT_DF_Z=1" "VXPT_DF_Z=0" "VXPT_DF_Z=1 VXPT_DF_XY_CTAS=384" "VXPT_DF_Z=1 VXPT_DF_XY_CTAS=148"; do
  echo "$env"; env $env timeout 120 python tools/df_probe.py 30 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print('  algo1', d['algo1'])"
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02j_launches_df.csv python tools/df_probe.py 3 > /dev/null 2>&1
grep -E "df_|pack" gpurun_out/r02j_launches_df.csv | head -4 | awk -F'","' '{print $5, $NF}' | cut -c1-200
for k in df_xy_dpx df_z_stream; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_r02j_$k python tools/df_probe.py 3 > gpurun_out/r02j_ncu_$k.log 2>&1
done
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py > gpurun_out/r02j_sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/r02j_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck python tools/sanitize_run.py > gpurun_out/r02j_sanitizer_racecheck.log 2>&1; tail -3 gpurun_out/r02j_sanitizer_racecheck.log
