# r02u: the scaling run as the driver does it, final code: N = 8, 4, 2 with --steps 20 --warmup 5; N = 8 with 200 steps
mkdir -p gpurun_out
run() { # N steps warmup tag extra...
  N=$1; S=$2; Wm=$3; T=$4; shift 4
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29730+RANDOM%200))"
  timeout 300 $TR bench.py --gpus $N --steps $S --warmup $Wm "$@" > gpurun_out/r02u_bench_n${N}_$T.json 2> gpurun_out/r02u_bench_n${N}_$T.err
  python - "$N" "$T" <<'PY'
import json, sys
n='gpurun_out/r02u_bench_n%s_%s.json' % (sys.argv[1], sys.argv[2])
try:
    d=json.loads([l for l in open(n) if l.startswith('{')][-1])
    print(n, round(d['value']), round(d['ms_per_step'],4), d['config']['submit'][:30], 'gathered', d.get('gathered_ok'), 'e2e', round(d['e2e']['value']), 'clocks', d['clocks']['samples'], d['clocks']['window'][:12])
except Exception as e: print(n, 'ERR', e)
PY
  grep -iE "error|fallback|failed|Traceback" gpurun_out/r02u_bench_n${N}_$T.err | head -3
}
run 8 20 5 driver
run 4 20 5 driver
run 2 20 5 driver
run 8 200 10 long
