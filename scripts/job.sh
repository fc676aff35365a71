set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cli.py -m gpu -q > gpurun_out/q_tests.log 2>&1
tail -n 30 gpurun_out/q_tests.log
UB200_REFREEZE=2 timeout 900 python -m pytest tests/test_cli.py -m gpu -q -k "ambiguous or sequential" > gpurun_out/q_tests2.log 2>&1
tail -n 8 gpurun_out/q_tests2.log
