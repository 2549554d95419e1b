# r03a: end-to-end loop with the frame call recorded into CUDA graphs (N = 1)
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-aux > gpurun_out/r03a_bench_n1.json 2> gpurun_out/r03a_bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r03a_bench_n1.json')); print(d['value'], d['ms_per_step'], d['e2e'])"; tail -3 gpurun_out/r03a_bench_n1.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-aux --emulate 8 > gpurun_out/r03a_bench_emu8.json 2> gpurun_out/r03a_bench_emu8.err
python -c "
import json; d=json.load(open('gpurun_out/r03a_bench_emu8.json')); print(d['value'], d['ms_per_step'], d['e2e'])"; tail -3 gpurun_out/r03a_bench_emu8.err
