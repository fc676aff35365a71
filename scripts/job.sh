set -x
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 200 python bench.py --no-extra > gpurun_out/y_bench.json 2> gpurun_out/y_bench.log
head -c 1500 gpurun_out/y_bench.json; echo; grep -E "flatten|spot check|cpu" gpurun_out/y_bench.log | head -5
