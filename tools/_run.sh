set -x
qb() { timeout 300 python tools/quick_bench.py llama-3-8b 128 512 0 > gpurun_out/qb_$1.log 2>&1; echo "=== $1"; grep -E 'xotb200:|"iter": 2|^qkv  |^o  |^gu |^down|^head|attention chain|step total|Error|error|barrier of' gpurun_out/qb_$1.log | cut -c1-330; }
XOTB200_PACE_GBS=7300 qb rel
XOTB200_PACE_GBS=7300 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_node_ring.py -m gpu -x -q 2>&1 | tail -5
