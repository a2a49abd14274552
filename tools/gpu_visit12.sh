#!/usr/bin/env bash
# Round-2 visit 12 (1 GPU): strong reject (all variants) + warp-convergent predicate (52, 54, 36) against 20 on one box.
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "[v12] drawprep + variants tests"; timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_variants.py tests/test_gpu_drawprep.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -6
for v in 20 52 54 36 20 52; do
  for w in C4 C2 target; do
    echo "[v12] variant $v workload $w"
    FYX_CULL_VARIANT=$v timeout 300 python bench.py --workload $w --no-c5 --no-parity --no-cpu-baseline --no-device-animation --steps 20 > $OUT/r02p_v${v}_$w.json 2> $OUT/r02p_v${v}_$w.err
    python - "$OUT/r02p_v${v}_$w.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    st = {k: (round(v["ms"], 4), round(v["frac"], 3)) for k, v in d["roofline"]["stages"].items()}
    print("   ms/frame", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), st)
except Exception as ex:
    print("   (no JSON line)", ex)
PY
  done
done
echo "[v12] done"
