#!/bin/bash
# multi-GPU visit: tools/gpu_scale.sh N  -> bench (MARS job, weak scaling) + retrieval sweep (strong scaling) on N GPUs
N=$1
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $RUN bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?" >> gpurun_out/bench_n$N.err
timeout 900 $RUN bench.py --gpus $N --workload sweep --steps 3 --warmup 1 > gpurun_out/sweep_n$N.json 2> gpurun_out/sweep_n$N.err; echo "rc=$?" >> gpurun_out/sweep_n$N.err
tail -n 2 gpurun_out/bench_n$N.err gpurun_out/sweep_n$N.err
tail -c 1200 gpurun_out/bench_n$N.json | head -c 1200; echo; tail -c 900 gpurun_out/sweep_n$N.json
