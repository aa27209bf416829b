#!/bin/bash
# ncu --set full of every kernel of ONE 882-tracklet head call (third pass: warm), plus the launch list of a bench step.
# usage: gpurun -- 'bash tools/ncu_head_r2.sh "split=4"'
mkdir -p gpurun_out
SPEC=${1:-split=4}
HV_TRACKLETS=882 HV_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"pool_tma_kernel|graph_kernel_tc|split_gemm_kernel|graph_mix_kernel|attn_kernel" -s 14 -c 7 -f \
    -o gpurun_out/prof_head_r2 python tools/head_variants.py 882 "$SPEC" > gpurun_out/ncu_head_r2.log 2>&1
tail -n 2 gpurun_out/ncu_head_r2.log | cut -c1-300
python tools/ncu_summary.py gpurun_out/prof_head_r2.ncu-rep > gpurun_out/ncu_head_r2.txt
grep -E "launch|Kernel Name|gpu__time_duration|dram__bytes|lts__throughput|tensor_cycles|warps_active|lts__t_sector_hit|xbar2l1tex" gpurun_out/ncu_head_r2.txt
