# r03i: ncu --set full captures of the final kernels + launch list of the bench command; bench with the reference-scale aux timings
mkdir -p gpurun_out
for k in gi_gen_trace0 gi_continue primary_kernel shadow_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_r03i_$k python tools/gi_probe.py 3 > gpurun_out/r03i_ncu_$k.log 2>&1
done
for k in df_xy_dpx df_z_dpx; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_r03i_$k python tools/df_probe.py 3 > gpurun_out/r03i_ncu_$k.log 2>&1
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r03i_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aux --no-graph > gpurun_out/r03i_bench_under_ncu.log 2>&1
grep -c "vxpt::" gpurun_out/r03i_launches_bench.csv
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r03i_bench_n1.json 2> gpurun_out/r03i_bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r03i_bench_n1.json')); print(d['value'], d['ms_per_step'], d['aux_passes'])"; tail -3 gpurun_out/r03i_bench_n1.err
