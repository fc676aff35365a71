set -x
mkdir -p gpurun_out
for cfg in "c4 0 32 1" "c4 0 96 3" "mid 2 32 1" "c3 1 96 3"; do
  UB200_PROFILE=1 timeout 300 python scripts/prof_roles.py $cfg >> gpurun_out/i_prof.log 2>&1
done
cat gpurun_out/i_prof.log
