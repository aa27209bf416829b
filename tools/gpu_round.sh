#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (+reference arm), re-ranking bench, ncu launch list, ncu full captures.
# Logs -> gpurun_out/.   usage: gpu_round.sh [nocapture]
mkdir -p gpurun_out
PT="python -m pytest -m gpu -q --no-header -p no:cacheprovider --timeout 180 --timeout-method=thread"
timeout 1200 $PT tests > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 python tools/rerank_bench.py > gpurun_out/rerank.json 2> gpurun_out/rerank.err
if [ "$1" != "nocapture" ]; then
# launch list of one whole step (13 head calls x 6 kernels + eval), after the warm-up steps
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 256 -c 100 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-fast-mode > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"pool_tma_kernel|pool_kernel|graph_kernel|split_gemm_kernel|attn_kernel" -s 12 -c 6 -f -o gpurun_out/prof_head \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-fast-mode > gpurun_out/ncu_head.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"split_gemm_kernel|rank_mars_kernel|rank_market_kernel|rank_mars_finish|rank_market_finish|split_planes" -s 8 -c 8 -f -o gpurun_out/prof_eval \
    python tools/run_eval_once.py 3 > gpurun_out/ncu_eval.log 2>&1
fi
tail -n 3 gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench.err gpurun_out/rerank.err
head -c 1500 gpurun_out/bench.json; echo; cat gpurun_out/rerank.json
