set -x
qb() { timeout 300 python tools/quick_bench.py llama-3-8b 128 512 0 > gpurun_out/qb_$1.log 2>&1; echo "=== $1"; grep -E '"iter": 2|^qkv  |^gu  |attention chain|step total|Error|error' gpurun_out/qb_$1.log | cut -c1-330; }
XOTB200_ATTN_L2PF=3 XOTB200_ATTN_DIRECT_MAX=1024 qb pf3_direct
XOTB200_ATTN_L2PF=7 XOTB200_ATTN_DIRECT_MAX=1024 qb pf7_direct
XOTB200_ATTN_L2PF=11 XOTB200_ATTN_DIRECT_MAX=1024 qb pf11_direct
XOTB200_ATTN_L2PF=8 XOTB200_ATTN_DIRECT_MAX=1024 qb pf8_direct
XOTB200_ATTN_L2PF=0 XOTB200_ATTN_DIRECT_MAX=1024 qb pf0_direct
XOTB200_ATTN_L2PF=11 XOTB200_ATTN_DIRECT_MAX=0 qb pf11_split
XOTB200_ATTN_L2PF=0 XOTB200_ATTN_DIRECT_MAX=0 qb pf0_split
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scheduler.py tests/test_gpu_node_ring.py -x -q 2>&1 | tail -3
