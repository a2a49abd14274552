#!/usr/bin/env bash
# Round-2 visit 4 (2 GPUs): single-GPU tests that changed, sub-forest kernel A/B, then the multi-GPU exchange.
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "[v4] GPU tests (all)"; timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $OUT/r02d_tests.log; tail -15 $OUT/r02d_tests.log
for sfm in 0 1; do
  for w in C3 C4; do
    echo "[v4] FYX_SUBFOREST=$sfm workload $w"
    FYX_SUBFOREST=$sfm timeout 300 python bench.py --workload $w --no-c5 --no-cpu-baseline --no-device-animation --steps 20 > $OUT/r02d_sf${sfm}_$w.json 2> $OUT/r02d_sf${sfm}_$w.err
    python - "$OUT/r02d_sf${sfm}_$w.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    st = {k: (round(v["ms"], 4), round(v["frac"], 3)) for k, v in d["roofline"]["stages"].items()}
    print("   ms/frame", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), st, "launches", d["gpu_launches"], "parity", d["parity"]["ok"])
except Exception as ex:
    print("   (no JSON line)", ex)
PY
  done
done
bash tools/gpu_visit_multi.sh 2 r02d full
