#!/bin/bash
# Captures behind profiles/rNN_*.md.  Run on the GPU box:  gpurun -- 'bash tools/profile_round.sh r01'
R=${1:-r01}; O=gpurun_out; mkdir -p $O
NCU="ncu --clock-control none"
# 1. launch list of the timed steps of the default bench command (one pass, durations only)
timeout 900 $NCU --metrics gpu__time_duration.sum --nvtx --nvtx-include "timed_steps/" -c 20000 --csv \
    --log-file $O/${R}_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-decode --no-cpu-baseline > $O/${R}_launches.log 2>&1
# 2. full-section captures of the hot kernels
FULL="$NCU --set full --import-source on -f"
timeout 300 $FULL -k regex:hessian_syrk_tc -s 2 -c 2 -o $O/${R}_syrk_tc python tools/prof_hessian.py > $O/${R}_ncu_syrk.log 2>&1
timeout 300 $FULL -k regex:woq_gemm_stream -s 30 -c 1 -o $O/${R}_gemv_stream_qkv python tools/prof_gemv.py 1 2 stream > $O/${R}_ncu_gemv1.log 2>&1
timeout 300 $FULL -k regex:woq_gemm_stream -s 40 -c 1 -o $O/${R}_gemv_stream_gateup python tools/prof_gemv.py 1 2 stream > $O/${R}_ncu_gemv2.log 2>&1
timeout 300 $FULL -k regex:woq_gemm_stream -s 50 -c 1 -o $O/${R}_gemv_stream_down python tools/prof_gemv.py 1 2 stream > $O/${R}_ncu_gemv3.log 2>&1
timeout 300 $FULL -k regex:woq_gemm_mma -s 40 -c 1 -o $O/${R}_gemv_cluster_M16_gateup python tools/prof_gemv.py 16 2 > $O/${R}_ncu_gemv4.log 2>&1
timeout 300 $FULL -k regex:gptq_subblock -s 8 -c 1 -o $O/${R}_gptq_subblock python tools/prof_gptq.py > $O/${R}_ncu_gptq1.log 2>&1
timeout 300 $FULL -k regex:gptq_lazy_update -s 8 -c 1 -o $O/${R}_gptq_lazy python tools/prof_gptq.py > $O/${R}_ncu_gptq2.log 2>&1
timeout 300 $FULL -k regex:quant_pack_kernel -c 1 -o $O/${R}_rtn_quant_pack python tools/prof_gptq.py > $O/${R}_ncu_rtn.log 2>&1
ls -la $O | grep ${R}_
tail -2 $O/${R}_launches.log
