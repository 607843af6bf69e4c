set -x
timeout 900 python -m pytest tests/test_gpu_tc_prefill.py tests/test_gpu_parity.py tests/test_gpu_full_depth.py -x -q 2>&1 | tail -5
for s in 512 1024 4096; do python tools/prefill_bench.py llama-3-8b $s 2>&1 | tail -1; XOTB200_TC_2SM=0 python tools/prefill_bench.py llama-3-8b $s 2>&1 | tail -1; done
