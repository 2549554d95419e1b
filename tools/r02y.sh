# r02y: final state — full GPU suite, smoke, bench as the driver runs it
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02y_pytest.log 2>&1; tail -3 gpurun_out/r02y_pytest.log
timeout 300 python -c "
import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02y_bench_n1.json 2> gpurun_out/r02y_bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02y_bench_n1.json')); print(d['value'], d['ms_per_step'], d['pass_ms'], d['roofline']['frac'], d['roofline'].get('issue_frac'), d['roofline_all']['df_build']['frac'], d['roofline_all']['df_build']['ms_per_launch'], d['e2e']['value'], d['gpu_launches'])"; tail -3 gpurun_out/r02y_bench_n1.err
