set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_cli.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/p_tests.log 2>&1
tail -n 40 gpurun_out/p_tests.log
