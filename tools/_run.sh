set -x
timeout 600 python -m pytest tests/test_gpu_node_ring.py tests/test_gpu_scheduler.py -m gpu -x -q 2>&1 | tail -4
bash tools/_run_multi.sh 2
