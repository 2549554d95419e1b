# r03j: xy sweep with a 16-bit tile (one DPX step per y step) vs the byte tile
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_df_step_field.py tests/test_per_frame_edits.py -m gpu -x -q -k "df or distance or step or edit" > gpurun_out/r03j_pytest.log 2>&1; tail -3 gpurun_out/r03j_pytest.log
for env in "VXPT_DF_XY=1" "VXPT_DF_XY=0" "VXPT_DF_XY=1" "VXPT_DF_XY=0" "VXPT_DF_XY=1 VXPT_PROBE_WORLD=city" "VXPT_DF_XY=0 VXPT_PROBE_WORLD=city"; do
  echo "$env"; env $env timeout 120 python tools/df_probe.py 40 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print('  algo1', d['algo1'])"
done
python tools/df_timeline.py
