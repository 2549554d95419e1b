# r02l: frames in flight at N = 8 / 4 under the driver's flags (--steps 20 --warmup 5); e2e with prepared frame calls
mkdir -p gpurun_out
run() { # N steps warmup tag extra...
  N=$1; S=$2; Wm=$3; T=$4; shift 4
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29730+RANDOM%200))"
  timeout 300 $TR bench.py --gpus $N --steps $S --warmup $Wm "$@" > gpurun_out/r02l_bench_n${N}_$T.json 2> gpurun_out/r02l_bench_n${N}_$T.err
  python - "$N" "$T" <<'PY'
import json, sys
n='gpurun_out/r02l_bench_n%s_%s.json' % (sys.argv[1], sys.argv[2])
try:
    d=json.loads([l for l in open(n) if l.startswith('{')][-1])
    print(n, round(d['value']), round(d['ms_per_step'],4), d['config']['submit'][:30], 'gathered', d.get('gathered_ok'), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'host', round(d['config']['host_submit_ms_per_step'],4))
except Exception as e: print(n, 'ERR', e)
PY
  grep -iE "error|fallback|failed|Traceback" gpurun_out/r02l_bench_n${N}_$T.err | head -3
}
run 8 20 5 p8
run 8 20 5 p12 --pipes 12
run 8 20 5 p16 --pipes 16
run 8 20 5 p20 --pipes 20
run 8 200 10 p16long --pipes 16
run 4 20 5 p8
run 4 20 5 p12 --pipes 12
