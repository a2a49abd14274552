#!/usr/bin/env bash
# Round-2 visit 13 (1 GPU): reworked bone fold (3 dependent round trips, cull by the lanes of one warp) and the fold on a side
# stream beside palette / skinning in asynchronous frames: tests, then FYX_SIDE_FOLD=0/1 on one box.
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "[v13] tests"; timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_drawprep.py tests/test_gpu_anim.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -6
for sfold in 0 1 0 1; do
  for w in C4 target C3; do
    echo "[v13] side fold $sfold workload $w"
    FYX_SIDE_FOLD=$sfold timeout 300 python bench.py --workload $w --no-c5 --no-cpu-baseline --no-device-animation --steps 20 > $OUT/r02q_sf${sfold}_$w.json 2> $OUT/r02q_sf${sfold}_$w.err
    python - "$OUT/r02q_sf${sfold}_$w.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    st = {k: (round(v["ms"], 4), round(v["frac"], 3)) for k, v in d["roofline"]["stages"].items()}
    print("   ms/frame", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), st, "parity", d.get("parity", {}).get("ok"))
except Exception as ex:
    print("   (no JSON line)", ex)
PY
  done
done
echo "[v13] done"
