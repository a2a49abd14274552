#!/usr/bin/env bash
# Round-2 visit 5 (1 GPU): deferred compaction (variant bit 3) against the default, variants test, ncu of the best.
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "[v5] variants + drawprep + parity tests"; timeout 1800 python -m pytest tests/test_gpu_variants.py tests/test_gpu_parity.py tests/test_gpu_drawprep.py -m gpu -q 2>&1 | tail -8
for v in 4 12 13 8; do
  for w in C4 C2 target C3; do
    echo "[v5] variant $v workload $w"
    FYX_CULL_VARIANT=$v timeout 300 python bench.py --workload $w --no-c5 --no-cpu-baseline --no-device-animation --steps 20 > $OUT/r02f_v${v}_$w.json 2> $OUT/r02f_v${v}_$w.err
    python - "$OUT/r02f_v${v}_$w.json" <<'PY'
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
for v in 12; do
  FYX_CULL_VARIANT=$v timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_update_level --launch-skip 3 -c 1 -o $OUT/r02f_full_c4_update_v$v \
      python bench.py --workload C4 --steps 2 --warmup 1 --no-c5 --no-parity --no-cpu-baseline --no-device-animation > $OUT/r02f_ncu_v$v.log 2>&1
  FYX_CULL_VARIANT=$v timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_compact_vis --launch-skip 1 -c 1 -o $OUT/r02f_full_c4_compact_v$v \
      python bench.py --workload C4 --steps 2 --warmup 1 --no-c5 --no-parity --no-cpu-baseline --no-device-animation > $OUT/r02f_ncu_c_v$v.log 2>&1
done
ls -la $OUT/r02f_*.ncu-rep
echo "[v5] done"
