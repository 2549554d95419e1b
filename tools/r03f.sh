# r03f: the driver's sequence on the final tree: GPU tests, smoke, reference arm, bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r03f_pytest.log 2>&1; tail -3 gpurun_out/r03f_pytest.log
timeout 300 python -c "
import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r03f_bench_reference_arm.json 2> gpurun_out/r03f_bench_reference_arm.err; cut -c1-200 gpurun_out/r03f_bench_reference_arm.json
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r03f_bench_n1.json 2> gpurun_out/r03f_bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r03f_bench_n1.json')); print(d['value'], d['ms_per_step'], d['pass_ms'], d['roofline']['frac'], d['roofline'].get('issue_frac'), d['roofline_all']['df_build']['frac'], d['e2e']['value'], d['e2e']['submit'][:30], d['gpu_launches'], d['clocks'], d['cpu_baseline']['value'])"; tail -3 gpurun_out/r03f_bench_n1.err
