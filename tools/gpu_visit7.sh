#!/usr/bin/env bash
# Round-2 visit 7 (1 GPU): variants 4 (40 registers), 20, 21 on one box.
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "[v7] drawprep + variants tests"; timeout 1800 python -m pytest tests/test_gpu_variants.py tests/test_gpu_drawprep.py -m gpu -q 2>&1 | tail -6
for v in 4 20 21 4 20 21; do
  for w in C4 C2 target; do
    echo "[v7] variant $v workload $w"
    FYX_CULL_VARIANT=$v timeout 300 python bench.py --workload $w --no-c5 --no-parity --no-cpu-baseline --no-device-animation --steps 20 > $OUT/r02i_v${v}_$w.json 2> $OUT/r02i_v${v}_$w.err
    python - "$OUT/r02i_v${v}_$w.json" <<'PY'
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
echo "[v7] done"
