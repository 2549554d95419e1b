# r04b: N = 2 under the driver's flags and the 2-GPU tests on the final library (GI continuation with even shares, register caps)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29771"
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r04b_bench_n2_driver.json 2> gpurun_out/r04b_bench_n2_driver.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r04b_bench_n2_driver.json') if l.startswith('{')][-1]); print(round(d['value']), d['ms_per_step'], d['config']['submit'][:30], d.get('gathered_ok'), round(d['e2e']['value']), d['e2e']['submit'][:20])"
grep -iE "error|fallback|failed|Traceback" gpurun_out/r04b_bench_n2_driver.err | head -3
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_mg_frame.py -m gpu -x -q 2>&1 | tail -3
