set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_cli.py -m gpu -x -q > gpurun_out/o_tests.log 2>&1
tail -n 6 gpurun_out/o_tests.log
timeout 600 python scripts/multi_check.py 2000000 4096 > gpurun_out/o_multi.log 2>&1
cat gpurun_out/o_multi.log
