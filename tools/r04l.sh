#!/bin/bash
# r04l: ncu --set full of df_xy_dpx after the y sweep moved to the lanes' high bytes (r04k)
mkdir -p gpurun_out
timeout 80 ncu --set full --clock-control none --import-source on -k regex:df_xy_dpx -s 3 -c 1 -f -o gpurun_out/prof_r04l_df_xy_dpx python tools/df_probe.py 3 > gpurun_out/r04l_ncu_df_xy_dpx.log 2>&1
tail -2 gpurun_out/r04l_ncu_df_xy_dpx.log
