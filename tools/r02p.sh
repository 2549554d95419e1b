# r02p: bench as the driver runs it (N = 1 here; N = 2 in the next call) after the NVML clock sampler, 16 frame indices, reflection wavefront
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02p_bench_n1.json 2> gpurun_out/r02p_bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02p_bench_n1.json')); print(d['value'], d['ms_per_step'], d['pass_ms'], d['roofline']['frac'], d['clocks'], d['e2e']['value'], d['cpu_baseline'])"; tail -3 gpurun_out/r02p_bench_n1.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02p_bench_reference_arm.json 2> gpurun_out/r02p_bench_reference_arm.err; cut -c1-600 gpurun_out/r02p_bench_reference_arm.json
timeout 300 python -c "
import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
