#!/usr/bin/env bash
# Round-2 visit 3 (1 GPU): all GPU tests (topology carry-over, blend shapes, bone blocks, static batches, kernel variants),
# then the FYX_UPDATE_ALL specialisation of the level kernel against the baseline.
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "[v3] all GPU tests"; timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -25 > $OUT/r02c_tests.log; tail -25 $OUT/r02c_tests.log
for v in 0 4 5; do
  for w in C4 C2 C3; do
    echo "[v3] variant $v workload $w"
    FYX_CULL_VARIANT=$v timeout 300 python bench.py --workload $w --no-c5 --no-cpu-baseline --no-device-animation --steps 20 > $OUT/r02c_v${v}_$w.json 2> $OUT/r02c_v${v}_$w.err
    python - "$OUT/r02c_v${v}_$w.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    st = {k: (round(v["ms"], 4), round(v["frac"], 3)) for k, v in d["roofline"]["stages"].items()}
    print("   ms/frame", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), st, "parity", d["parity"]["ok"])
except Exception as ex:
    print("   (no JSON line)", ex)
PY
  done
done
echo "[v3] done"
