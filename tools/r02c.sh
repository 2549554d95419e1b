mkdir -p gpurun_out
for k in gi_continue gi_gen_trace0; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_r02c_$k python tools/gi_probe.py 2 > gpurun_out/r02c_ncu_$k.log 2>&1
done
ls -la gpurun_out | grep r02c
