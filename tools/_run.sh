set -x
bash tools/_run_multi.sh 2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 4 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; tail -c 2500 gpurun_out/bench_ref_n2.json; tail -3 gpurun_out/bench_ref_n2.err
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_node_ring.py -x -q 2>&1 | tail -3
