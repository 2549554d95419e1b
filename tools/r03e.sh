# r03e: BASELINE configs[3] (2160p, 4-spp GI) and configs[4] (per-frame edit + rebuild) on 8 GPUs
mkdir -p gpurun_out
run() { # N config steps tag
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29730+RANDOM%200))"
  timeout 400 $TR bench.py --gpus $1 --config $2 --steps $3 --warmup 5 > gpurun_out/r03e_bench_config$2_n$1.json 2> gpurun_out/r03e_bench_config$2_n$1.err
  python - "$1" "$2" <<'PY'
import json, sys
n='gpurun_out/r03e_bench_config%s_n%s.json' % (sys.argv[2], sys.argv[1])
try:
    d=json.loads([l for l in open(n) if l.startswith('{')][-1])
    print(n, round(d['value']), round(d['ms_per_step'],4), d['config']['submit'][:24], 'gathered', d.get('gathered_ok'), 'e2e', round(d['e2e']['value']), d.get('rebuild'))
except Exception as e: print(n, 'ERR', e)
PY
  grep -iE "error|fallback|failed|Traceback" gpurun_out/r03e_bench_config$2_n$1.err | head -3
}
run 8 4 50 
run 8 5 50
