# r02e: 2-GPU validation — gathered-frame tests, bench with the driver's flags (graph path at N=2, gathered_ok), N=1 bench after the GI default change
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02e_pytest_multi.log 2>&1; tail -3 gpurun_out/r02e_pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29721 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02e_bench_n2.json 2> gpurun_out/r02e_bench_n2.err
python - <<'PY'
import json
for n in ('gpurun_out/r02e_bench_n2.json',):
    try:
        d=json.loads([l for l in open(n) if l.startswith('{')][-1])
        print(n, d['value'], d['ms_per_step'], d['config']['submit'][:60], d.get('gathered_ok'), d['e2e']['value'], d['config']['host_submit_ms_per_step'])
    except Exception as e: print(n, 'ERR', e)
PY
grep -iE "error|fallback|failed|Traceback" gpurun_out/r02e_bench_n2.err | head -5
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02e_bench_n1.json')); print(d['value'], d['ms_per_step'], d['pass_ms'], d['roofline']['frac'], d['e2e']['value'])"
