# r02t: frame pipes on one GPU
mkdir -p gpurun_out
for p in 2 3 4 6; do
  timeout 300 python bench.py --pipes $p --steps 20 --warmup 5 --no-cpu-baseline --no-aux > gpurun_out/r02t_bench_p$p.json 2> gpurun_out/r02t_bench_p$p.err
  python -c "
import json; d=json.load(open('gpurun_out/r02t_bench_p$p.json')); print('pipes $p steps 20:', round(d['value']), d['ms_per_step'], d['roofline_all']['df_build']['ms_per_launch'])"
  timeout 300 python bench.py --pipes $p --steps 200 --warmup 10 --no-cpu-baseline --no-aux > gpurun_out/r02t_bench_p${p}_long.json 2> gpurun_out/r02t_bench_p${p}_long.err
  python -c "
import json; d=json.load(open('gpurun_out/r02t_bench_p${p}_long.json')); print('pipes $p steps 200:', round(d['value']), d['ms_per_step'])"
done
