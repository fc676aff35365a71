set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
timeout 200 python bench.py --no-extra > gpurun_out/w_bench.json 2> gpurun_out/w_bench.log
head -c 1200 gpurun_out/w_bench.json; echo; tail -n 3 gpurun_out/w_bench.log
