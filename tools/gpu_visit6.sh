#!/usr/bin/env bash
# Round-2 visit 6 (1 GPU): 32-register builds of the level kernel (variants 20, 28) against the default (4); new tests.
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "[v6] tests: variants + drawprep"; timeout 1800 python -m pytest tests/test_gpu_variants.py tests/test_gpu_drawprep.py -m gpu -q 2>&1 | tail -8
for v in 4 20 28; do
  for w in C4 C2 target; do
    echo "[v6] variant $v workload $w"
    FYX_CULL_VARIANT=$v timeout 300 python bench.py --workload $w --no-c5 --no-cpu-baseline --no-device-animation --steps 20 > $OUT/r02h_v${v}_$w.json 2> $OUT/r02h_v${v}_$w.err
    python - "$OUT/r02h_v${v}_$w.json" <<'PY'
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
FYX_CULL_VARIANT=20 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_update_level --launch-skip 3 -c 1 -o $OUT/r02h_full_c4_update_v20 \
    python bench.py --workload C4 --steps 2 --warmup 1 --no-c5 --no-parity --no-cpu-baseline --no-device-animation > $OUT/r02h_ncu_v20.log 2>&1
ls -la $OUT/r02h_*.ncu-rep
echo "[v6] done"
