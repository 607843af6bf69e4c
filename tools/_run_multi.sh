# multi-GPU validation: bench at N (both arms), default flags like the driver
set -x
N=$1; shift
PORT=$((29500 + N))
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 3 --warmup 3 "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -c 4000 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
