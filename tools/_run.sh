set -x
nproc; free -g | head -2
python -m pytest tests/test_gpu_full_depth.py -x -q -s 2>&1 | tail -15
for v in 8_4 6_3 4_2 8_2; do
  echo "=== variant $v"
  XOTB200_LIB=$PWD/xotorch_b200/lib/ab_$v.so python tools/quick_bench.py llama-3-8b 128 512 0 2>&1 | grep -E '"iter": 2|^qkv  |attention chain|step total'
done
