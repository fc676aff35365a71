set -x
timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r_launch_bench.log 2>&1
tail -c 300 gpurun_out/r_launch_bench.log; wc -l gpurun_out/r_launches.csv
