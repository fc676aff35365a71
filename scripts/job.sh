set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/u_tests.log 2>&1
tail -n 10 gpurun_out/u_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/u_bench_n1.json 2> gpurun_out/u_bench_n1.log
head -c 1500 gpurun_out/u_bench_n1.json; echo
grep "extra\|spot check" gpurun_out/u_bench_n1.log | cut -c1-230
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/u_bench_ref.json 2> gpurun_out/u_bench_ref.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/u_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_score4 -s 1 -c 1 -f -o gpurun_out/r02_k4_c4_b32 python scripts/one_launch.py c4 0 32 32 2 1 > gpurun_out/u_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_score4 -s 1 -c 1 -f -o gpurun_out/r02_k4_c4_b96 python scripts/one_launch.py c4 0 96 96 2 3 > gpurun_out/u_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_score4 -s 1 -c 1 -f -o gpurun_out/r02_k4_c3_b256 python scripts/one_launch.py c3 1 256 256 2 0 > gpurun_out/u_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_score4 -s 1 -c 1 -f -o gpurun_out/r02_k4_c4_leaf python scripts/one_launch.py c4 1 32 32 2 1 > gpurun_out/u_ncu4.log 2>&1
for tool in memcheck synccheck racecheck; do
  UB200_MIN_TILE=300 timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_small.py > gpurun_out/r_san_$tool.log 2>&1
  tail -n 2 gpurun_out/r_san_$tool.log
done
python scripts/fs_bench.py 200000 2000 2>&1 | tail -2 > gpurun_out/u_fs.log; cat gpurun_out/u_fs.log
