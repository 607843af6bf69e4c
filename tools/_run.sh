set -x
timeout 600 python -m pytest tests/test_gpu_tc_prefill.py -x -q 2>&1 | tail -3
for s in 512 1024 4096; do python tools/prefill_bench.py llama-3-8b $s 2>&1 | tail -1; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_prefill1024_launches.csv python tools/prefill_bench.py llama-3-8b 1024 > gpurun_out/ncu_pf.log 2>&1; python tools/ncu_summarize.py launches gpurun_out/r2_prefill1024_launches.csv | head -8
