# r02g: full GPU suite after the jittered reflection read; bench N=1 with the auxiliary pass timings; ncu of the reflection kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02g_pytest.log 2>&1; tail -4 gpurun_out/r02g_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02g_bench_n1.json 2> gpurun_out/r02g_bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02g_bench_n1.json')); print(d['value'], d['ms_per_step'], d['pass_ms'], d['roofline']['frac'], d['e2e']['value'], d['aux_passes'])"; tail -3 gpurun_out/r02g_bench_n1.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:reflection_kernel -s 2 -c 1 -f -o gpurun_out/prof_r02g_reflection python tools/denoise_probe.py 2 > gpurun_out/r02g_ncu_reflection.log 2>&1; tail -2 gpurun_out/r02g_ncu_reflection.log
