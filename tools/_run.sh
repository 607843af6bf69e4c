set -x
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tc_ -c 700 --csv --log-file gpurun_out/r2b_prefill128_launches.csv python tools/prefill_bench.py llama-3-8b 128 > gpurun_out/ncu_prefill128.log 2>&1
tail -3 gpurun_out/ncu_prefill128.log
wc -l gpurun_out/r2b_prefill128_launches.csv
