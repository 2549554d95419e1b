N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "2" ]; then
timeout 120 $TR --master-port 29701 tools/p2p_check.py p2p texel > gpurun_out/p2p_check.log 2>&1; tail -1 gpurun_out/p2p_check.log
timeout 120 $TR --master-port 29702 tools/p2p_check.py p2pcopy texel >> gpurun_out/p2p_check.log 2>&1; tail -1 gpurun_out/p2p_check.log
fi
i=0
for extra in "${@:2}"; do
  i=$((i+1))
  timeout 150 $TR --master-port $((29710+i)) bench.py --gpus $N --steps 300 --warmup 10 $extra > gpurun_out/bench_g${N}_$i.json 2> gpurun_out/bench_g${N}_$i.err
  echo "run $i [$extra] rc=$?"; grep '^{' gpurun_out/bench_g${N}_$i.json | python tools/show.py "$extra"; grep -iE "error|fallback|Traceback" gpurun_out/bench_g${N}_$i.err | head -3
done
