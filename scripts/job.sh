set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_score4 -s 1 -c 1 -f -o gpurun_out/m_c4 python scripts/one_launch.py c4 0 32 32 2 1 > gpurun_out/m_c4_ncu.log 2>&1
tail -n 3 gpurun_out/m_c4_ncu.log
