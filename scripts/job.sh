set -x
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_cli.py -m gpu -x -q > gpurun_out/t_tests.log 2>&1
tail -n 4 gpurun_out/t_tests.log
python - <<'PY'
import sys, time
sys.path.insert(0, ".")
from usher_b200 import capi
s = capi.Synth(10_000_000, 30.0, 30000, 0, 20260929)
m = capi.Mat.from_flat_struct(s.flat)
for fam in (0, 1, 2):
    sp, sc, _ = s.samples(256, fam, 11)
    S = m.upload(sp, sc)
    for flags in (0, 2):
        S.place(flags); t = time.time(); S.place(flags); S.place(flags); dt = (time.time() - t) / 2
        print(f"c4 fam={fam} flags={flags}: {dt*1e3:.1f} ms per 256 samples -> {256/dt:.0f} placements/s", flush=True)
    S.close()
PY
