set -x
qb() { timeout 300 python tools/quick_bench.py llama-3-8b 128 512 0 > gpurun_out/qb_$1.log 2>&1; echo "=== $1"; grep -E 'xotb200:|"iter": 2|^qkv  |attention chain|step total|Error|error' gpurun_out/qb_$1.log | cut -c1-330; }
export XOTB200_LIB=$PWD/xotorch_b200/lib/ab_direct6.so
XOTB200_KV_PERSIST=1 XOTB200_ATTN_DIRECT_MAX=0 qb d6_split_persist
XOTB200_KV_PERSIST=1 XOTB200_ATTN_DIRECT_MAX=1024 qb d6_direct_persist
XOTB200_KV_PERSIST=0 XOTB200_ATTN_DIRECT_MAX=1024 qb d6_direct
