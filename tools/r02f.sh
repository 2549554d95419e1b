# r02f: the scaling run as the driver does it (--steps 20 --warmup 5) at N = 8 and 4, plus a longer run at 8
mkdir -p gpurun_out
run() { # N steps warmup tag
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29730+$1))"
  timeout 300 $TR bench.py --gpus $1 --steps $2 --warmup $3 > gpurun_out/r02f_bench_n$1_$4.json 2> gpurun_out/r02f_bench_n$1_$4.err
  python - "$1" "$4" <<'PY'
import json, sys
n='gpurun_out/r02f_bench_n%s_%s.json' % (sys.argv[1], sys.argv[2])
try:
    d=json.loads([l for l in open(n) if l.startswith('{')][-1])
    print(n, round(d['value']), round(d['ms_per_step'],4), d['config']['submit'][:40], 'gathered', d.get('gathered_ok'), 'e2e', round(d['e2e']['value']), 'host', round(d['config']['host_submit_ms_per_step'],4))
except Exception as e: print(n, 'ERR', e)
PY
  grep -iE "error|fallback|failed|Traceback" gpurun_out/r02f_bench_n$1_$4.err | head -3
}
run 8 20 5 driver
run 4 20 5 driver
run 8 200 10 long
