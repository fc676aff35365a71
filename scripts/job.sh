set -x
python scripts/one_launch.py mid 0 256 32 > gpurun_out/a_mid0.log 2>&1
python scripts/one_launch.py mid 1 256 32 > gpurun_out/a_mid1.log 2>&1
python scripts/one_launch.py mid 2 256 32 > gpurun_out/a_mid2.log 2>&1
python scripts/one_launch.py c3 1 256 256 > gpurun_out/a_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_score3 -s 2 -c 1 -f -o gpurun_out/a_mid1 python scripts/one_launch.py mid 1 32 32 2 > gpurun_out/a_mid1_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_score3 -s 2 -c 1 -f -o gpurun_out/a_mid2 python scripts/one_launch.py mid 2 32 32 2 > gpurun_out/a_mid2_ncu.log 2>&1
tail -3 gpurun_out/a_mid0.log gpurun_out/a_mid1.log gpurun_out/a_mid2.log gpurun_out/a_c3.log
