set -x
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_prefill1024_launches.csv python tools/prefill_bench.py llama-3-8b 1024 > gpurun_out/ncu_pf.log 2>&1; python tools/ncu_summarize.py launches gpurun_out/r2_prefill1024_launches.csv > gpurun_out/r2_prefill1024_launch_list_summary.txt; cat gpurun_out/r2_prefill1024_launch_list_summary.txt
timeout 600 ncu --set full --clock-control none -k regex:tc_gemm_2sm -s 200 -c 4 -o gpurun_out/r2_gemm2sm python tools/prefill_bench.py llama-3-8b 1024 > gpurun_out/ncu_2sm.log 2>&1; tail -2 gpurun_out/ncu_2sm.log
ncu -i gpurun_out/r2_gemm2sm.ncu-rep --page raw --csv > gpurun_out/r2_gemm2sm_raw.csv 2>/dev/null; python tools/ncu_summarize.py full gpurun_out/r2_gemm2sm_raw.csv
