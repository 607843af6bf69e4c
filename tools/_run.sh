set -x
qb() { timeout 300 python tools/quick_bench.py llama-3-8b 128 512 0 > gpurun_out/qb_$1.log 2>&1; echo "=== $1"; grep -E '"iter": 2|^qkv  |^o  |^gu  |attention chain|step total|Error|error' gpurun_out/qb_$1.log | cut -c1-330; }
export XOTB200_LIB=$PWD/xotorch_b200/lib/ab_l2a.so
XOTB200_L2_AHEAD=0 qb l2a_0
XOTB200_L2_AHEAD=4 qb l2a_4
XOTB200_L2_AHEAD=12 qb l2a_12
XOTB200_L2_AHEAD=24 qb l2a_24
