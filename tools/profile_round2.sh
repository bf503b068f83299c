#!/bin/bash
# Captures behind profiles/r02_*.md.  Run on the GPU box:  gpurun -- 'bash tools/profile_round2.sh'
R=r02; O=gpurun_out; mkdir -p $O
NCU="ncu --clock-control none"
# 1. launch list of one timed step of the default bench command (durations only)
timeout 900 $NCU --metrics gpu__time_duration.sum --nvtx --nvtx-include "timed_steps/" -c 20000 --csv \
    --log-file $O/${R}_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-decode --no-cpu-baseline --no-wall > $O/${R}_launches.log 2>&1
# 2. full-section captures of the hot / new kernels
FULL="$NCU --set full --import-source on -f"
timeout 300 $FULL -k regex:hessian_syrk_tc2 -s 1 -c 1 -o $O/${R}_syrk_tc_4096 python tools/prof_r02.py syrk > $O/${R}_ncu1.log 2>&1
timeout 300 $FULL -k regex:hessian_syrk_tc2 -s 4 -c 1 -o $O/${R}_syrk_tc_11008 python tools/prof_r02.py syrk > $O/${R}_ncu2.log 2>&1
timeout 300 $FULL -k regex:potrf_inv_tile -s 3 -c 1 -o $O/${R}_k2_diag_tile python tools/prof_r02.py k2 4096 > $O/${R}_ncu3.log 2>&1
timeout 300 $FULL -k regex:chol_trail -s 2 -c 1 -o $O/${R}_k2_trail python tools/prof_r02.py k2 4096 > $O/${R}_ncu4.log 2>&1
timeout 300 $FULL -k regex:woq_gemm_tc -s 4 -c 1 -o $O/${R}_woq_tc_M16 python tools/prof_r02.py woq_tc 16 > $O/${R}_ncu5.log 2>&1
timeout 300 $FULL -k regex:w8a8_gemm -s 4 -c 1 -o $O/${R}_w8a8_M2048 python tools/prof_r02.py w8a8 2048 > $O/${R}_ncu7.log 2>&1
ls -la $O | grep ${R}_
tail -2 $O/${R}_launches.log
