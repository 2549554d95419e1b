# r03g: z-sweep carries: two-level (shuffles + one barrier) vs the round-1 serial scan
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_df_step_field.py tests/test_per_frame_edits.py -m gpu -x -q -k "df or distance or step or edit" > gpurun_out/r03g_pytest.log 2>&1; tail -3 gpurun_out/r03g_pytest.log
for lib in libvxpt.so libvxpt_sc.so libvxpt.so libvxpt_sc.so; do
  echo "$lib"; VXPT_LIB=$lib timeout 120 python tools/df_probe.py 40 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print('  algo1', d['algo1'])"
done
python tools/df_timeline.py
