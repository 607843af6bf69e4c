set -x
qb() { timeout 300 python tools/quick_bench.py llama-3-8b 128 512 0 > gpurun_out/qb_$1.log 2>&1; echo "=== $1"; grep -E '"iter": 2|^qkv  |attention chain|step total|Error|error' gpurun_out/qb_$1.log | cut -c1-330; }
qb merge8
XOTB200_NSPLIT_MIN=8 qb merge8_ns8
XOTB200_NSPLIT_MIN=6 qb merge8_ns6
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_depth.py -x -q 2>&1 | tail -3
