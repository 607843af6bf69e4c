set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err; tail -c 600 gpurun_out/bench_n1_final.json; tail -2 gpurun_out/bench_n1_final.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:shard_step -s 1 -c 1 -f -o gpurun_out/r2b_mega python tools/quick_bench.py llama-3-8b 128 48 > gpurun_out/ncu_mega.log 2>&1; tail -2 gpurun_out/ncu_mega.log
ls -la gpurun_out/r2b_mega.ncu-rep
