#!/usr/bin/env bash
# Round-2 visit 9 (1 GPU): scatter of the changed transforms on the copy stream (default) against the main-stream scatter; smoke().
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "[v9] smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "[v9] tests"; timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_anim.py -m gpu -q 2>&1 | tail -4
for side in 1 0 1 0; do
  for w in C4 C3; do
    echo "[v9] FYX_SIDE_SCATTER=$side workload $w"
    FYX_SIDE_SCATTER=$side timeout 300 python bench.py --workload $w --no-c5 --no-cpu-baseline --no-device-animation --steps 20 > $OUT/r02k_side${side}_$w.json 2> $OUT/r02k_side${side}_$w.err
    python - "$OUT/r02k_side${side}_$w.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("   ms/frame", round(d["ms_per_step"], 4), "e2e pipelined", round(d["e2e"]["ms_per_step_pipelined"], 4), "sync", round(d["e2e"]["ms_per_step_synchronous"], 4),
          "static+skeletons", round(d["modes"]["static_plus_skeletons_e2e_ms_per_step"], 4), "parity", d["parity"]["ok"])
except Exception as ex:
    print("   (no JSON line)", ex)
PY
  done
done
echo "[v9] done"
