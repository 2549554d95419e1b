# r02b: GPU parity of the staged gi_continue + LUT trigonometry, GI timing variants, a short bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; tail -4 gpurun_out/r02b_pytest.log
for ctas in 3 4 6; do VXPT_GI_CTAS=$ctas timeout 120 python tools/gi_probe.py 20 >> gpurun_out/r02b_gi_probe.jsonl 2>> gpurun_out/r02b_gi_probe.err; done
for ctas in 4 6; do VXPT_LIB=libvxpt_exp1.so VXPT_GI_CTAS=$ctas timeout 120 python tools/gi_probe.py 20 >> gpurun_out/r02b_gi_probe.jsonl 2>> gpurun_out/r02b_gi_probe.err; done
VXPT_GI_SORT=2 timeout 120 python tools/gi_probe.py 20 >> gpurun_out/r02b_gi_probe.jsonl 2>> gpurun_out/r02b_gi_probe.err
cat gpurun_out/r02b_gi_probe.jsonl
timeout 300 python bench.py --steps 100 --warmup 5 > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; cat gpurun_out/r02b_bench_n1.json | cut -c1-1500; tail -3 gpurun_out/r02b_bench_n1.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02b_launches_gi.csv python tools/gi_probe.py 2 > /dev/null 2>&1
grep -E "gi_|primary|shadow" gpurun_out/r02b_launches_gi.csv | tail -12 | cut -c1-300
