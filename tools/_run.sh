set -x
for cfg in 64,8 32,8 32,4 64,4 128,6; do echo "TC_SMALL=$cfg"; XOTB200_TC_SMALL=$cfg python tools/prefill_bench.py llama-3-8b 128 2>&1 | tail -1; done
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
