#!/bin/bash
# One command per multi-GPU configuration of BASELINE.json, launched the way the driver launches bench.py (one rank per GPU):
#   bash tools/bench_multi.sh 2                                        # config 2: Llama-3-8B, 2 shards
#   bash tools/bench_multi.sh 8                                        # config 3: Llama-3-8B, 8 shards
#   bash tools/bench_multi.sh 8 --model llama-3-70b --prompt 4096      # config 4: Llama-3-70B fp16, 8 shards, 4k prompt
#   bash tools/bench_multi.sh 4 --model mixtral-8x7b                   # config 5: Mixtral-8x7B, 4 shards
#   bash tools/bench_multi.sh 2 --impl reference                       # the reference arm (CPU engine + reference gRPC hop)
# The JSON line lands in gpurun_out/bench_n<N>.json (copy the ones you want judged to profiles/).
set -x
N=$1; shift
PORT=$((29500 + N))
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 3 --warmup 3 "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 4000 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
