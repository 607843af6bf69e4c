set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 2 --warmup 1 --model mixtral-8x7b > gpurun_out/bench_mixtral_n4.json 2> gpurun_out/bench_mixtral_n4.err; tail -c 3000 gpurun_out/bench_mixtral_n4.json; tail -3 gpurun_out/bench_mixtral_n4.err
timeout 600 python -m pytest tests/test_gpu_node_ring.py tests/test_gpu_multi.py tests/test_gpu_scheduler.py -x -q 2>&1 | tail -3
