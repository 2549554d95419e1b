# r02n: reflection pass re-queued (refl_gen_trace + refl_shade) vs one thread per pixel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_material_pass.py tests/test_z_material_extras.py tests/test_mg_frame.py -m gpu -x -q -k "reflection or reflections or feeds or halo or mg_frame" > gpurun_out/r02n_pytest.log 2>&1; tail -4 gpurun_out/r02n_pytest.log
timeout 200 python tools/denoise_probe.py 10 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print({k: round(v['ms'],4) for k,v in d['passes'].items()})"
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum --clock-control none -k regex:refl_ -c 8 --csv --log-file gpurun_out/r02n_launches_refl.csv python tools/denoise_probe.py 2 > /dev/null 2>&1
grep -E "refl_" gpurun_out/r02n_launches_refl.csv | tail -6 | awk -F'","' '{print $5, $(NF-2), $NF}' | cut -c1-220
for k in refl_gen_trace refl_shade; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_r02n_$k python tools/denoise_probe.py 2 > gpurun_out/r02n_ncu_$k.log 2>&1
done
