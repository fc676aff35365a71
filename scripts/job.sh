set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or sharing or tiny or random_vs_port" > gpurun_out/s_tests.log 2>&1
tail -n 4 gpurun_out/s_tests.log
python scripts/quick_bench2.py c4 0,1 32:1,96:3 2>&1 | tail -4
python scripts/quick_bench2.py mid,c3 0,2 32:1,256:3 2>&1 | grep -v "^\["
