#!/bin/bash
# One GPU-box visit for the committed round-2 profiles: ncu --set full of one head call and of the evaluation kernels,
# the launch list (gpu__time_duration) of one bench step.   usage: gpurun -- 'bash tools/gpu_profiles_r2.sh'
mkdir -p gpurun_out
HV_TRACKLETS=882 HV_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"pool_tma_kernel|graph_kernel_tc|gemm_kernel|graph_mix_kernel|attn_kernel" -s 14 -c 7 -f \
    -o gpurun_out/prof_head_final python tools/head_variants.py 882 "split=4" > gpurun_out/ncu_head_final.log 2>&1
tail -n 2 gpurun_out/ncu_head_final.log | cut -c1-300
python tools/ncu_summary.py gpurun_out/prof_head_final.ncu-rep > gpurun_out/ncu_head_final.txt
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:"gemm_kernel|rank_mars_kernel|rank_market_kernel|rank_mars_finish|rank_market_finish|split_planes" -s 8 -c 8 -f \
    -o gpurun_out/prof_eval_final python tools/run_eval_once.py 3 > gpurun_out/ncu_eval_final.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_eval_final.ncu-rep > gpurun_out/ncu_eval_final.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 120 --csv --log-file gpurun_out/launches_bench_final.csv \
    python bench.py --steps 1 --warmup 3 --quick > gpurun_out/ncu_launch.log 2>&1
grep -E "launch|Kernel Name|gpu__time_duration|dram__bytes|tensor_cycles" gpurun_out/ncu_head_final.txt | head -80
tail -5 gpurun_out/launches_bench_final.csv | cut -c1-200
