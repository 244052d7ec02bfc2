#!/bin/bash
# Round profile capture (run under gpurun, one GPU): launch list, ncu --set full of the top kernels, the bench line.
set -x
R=${ROUND:-r2}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/prof_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k 'regex:k_chol_rs|k_schur_tiles|k_backsolve_w|k_front_syrk|k_proj_obs|k_zmat|k_proj_pose|k_lm_backsub_obs|k_schur_rhs|k_preintegrate' -s 30 -c 16 \
    -o gpurun_out/${R}_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/prof_full.log 2>&1
FG_CHOL_TRACE=gpurun_out/${R}_chol_trace.txt timeout 300 python bench.py --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/prof_trace.log 2>&1
timeout 600 python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err
tail -c 400 gpurun_out/${R}_bench_n1.json
