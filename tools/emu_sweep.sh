for cfg in "8 2" "8 4" "8 8" "8 12" "4 4" "4 8" "2 4"; do
  set -- $cfg
  timeout 100 python bench.py --steps 400 --warmup 10 --no-cpu-baseline --emulate $1 --pipes $2 > gpurun_out/bench_emu_$1_$2.json 2> gpurun_out/bench_emu_$1_$2.err
  grep '^{' gpurun_out/bench_emu_$1_$2.json | python tools/show.py "emulate=$1 pipes=$2"
done
