# round-end measurement on one B200: GPU tests, default bench (+ CPU baseline), reference arm, GI-mode comparison, ncu launch list and
# full-set captures of every kernel on the path.  Outputs under gpurun_out/.
R=${1:-r01g}
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$R.log 2>&1; tail -2 gpurun_out/pytest_gpu_$R.log
timeout 300 python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; python tools/show.py default < gpurun_out/bench_$R.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$R.json 2>> gpurun_out/bench_$R.err; cut -c1-220 gpurun_out/bench_ref_$R.json


timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --pipes 1 > gpurun_out/ncu_bench_$R.log 2>&1
for k in primary_kernel shadow_kernel gi_gen_trace0 gi_continue df_xy_dpx df_z_dpx pack_steps; do
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/prof_${R}_$k python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --pipes 1 >> gpurun_out/ncu_bench_$R.log 2>&1
done
ls gpurun_out | grep $R | tr '\n' ' '
