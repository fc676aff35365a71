set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_score4 -s 1 -c 1 -f -o gpurun_out/e_mid1 python scripts/one_launch.py mid 1 32 32 2 1 > gpurun_out/e_mid1_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_score4 -s 1 -c 1 -f -o gpurun_out/e_mid0nc3 python scripts/one_launch.py mid 0 96 96 2 3 > gpurun_out/e_mid0nc3_ncu.log 2>&1
tail -n 4 gpurun_out/e_mid1_ncu.log gpurun_out/e_mid0nc3_ncu.log
