timeout 170 python scripts/variants.py c4 0 256:32:1,1152:96:3 base acc2 nofence tw acc2_nofence_tw base 2>&1 | tee gpurun_out/x_variants.log | tail -14
