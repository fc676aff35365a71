set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --pass-samples 32 --no-cpu-baseline --no-extra > gpurun_out/n_bench32.json 2> gpurun_out/n_bench32.log
timeout 600 python bench.py --steps 5 --warmup 3 --pass-samples 96 --no-cpu-baseline --no-extra > gpurun_out/n_bench96.json 2> gpurun_out/n_bench96.log
timeout 600 python bench.py --steps 5 --warmup 3 --pass-samples 64 --no-cpu-baseline --no-extra > gpurun_out/n_bench64.json 2> gpurun_out/n_bench64.log
python - <<'PY'
import json
for f in ("n_bench32","n_bench64","n_bench96"):
    try:
        j=json.load(open(f"gpurun_out/{f}.json")); r=j["roofline"]
        print(f, "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "us/launch", round(r["us_per_launch"],1), "frac", round(r["frac"],3), "share", round(r["share_of_step"],3), "launches", r["launches"])
    except Exception as e: print(f, "failed", e)
PY
