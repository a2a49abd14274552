#!/usr/bin/env bash
# Round-2 visit 2 (1 GPU): cull variants A/B (warp-union pre-reject, warp-wide compaction) on C4 / C2 / target, parity on.
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "[v2] all GPU tests (new cull, topology carry-over, kernel variants)"; timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
for v in 0 1 2 3; do
  for w in C4 C2; do
    echo "[v2] variant $v workload $w"
    FYX_CULL_VARIANT=$v timeout 300 python bench.py --workload $w --no-c5 --no-cpu-baseline --no-device-animation --steps 20 > $OUT/r02b_v${v}_$w.json 2> $OUT/r02b_v${v}_$w.err
    python - "$OUT/r02b_v${v}_$w.json" <<'PY'
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
echo "[v2] default variant: target, C3"
for w in target C3; do
  timeout 300 python bench.py --workload $w --no-c5 --no-cpu-baseline --no-device-animation --steps 20 > $OUT/r02b_def_$w.json 2> $OUT/r02b_def_$w.err
  python - "$OUT/r02b_def_$w.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    st = {k: (round(v["ms"], 4), round(v["frac"], 3)) for k, v in d["roofline"]["stages"].items()}
    print("   ms/frame", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), st, "parity", d["parity"]["ok"])
except Exception as ex:
    print("   (no JSON line)", ex)
PY
done
echo "[v2] skin variants on C4"
for v in tma2 tma3; do
  FYX_SKIN_VARIANT=$v timeout 300 python bench.py --workload C4 --no-c5 --no-cpu-baseline --no-device-animation --steps 20 > $OUT/r02b_skin_${v}.json 2> $OUT/r02b_skin_${v}.err
  python - "$OUT/r02b_skin_${v}.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    st = {k: (round(v["ms"], 4), round(v["frac"], 3)) for k, v in d["roofline"]["stages"].items()}
    print("   ms/frame", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), st, "parity", d["parity"]["ok"])
except Exception as ex:
    print("   (no JSON line)", ex)
PY
done
for v in ldg tma2; do
  FYX_SKIN_VARIANT=$v timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_skin --launch-skip 2 -c 1 -o $OUT/r02b_full_skin_$v \
      python bench.py --workload target --steps 2 --warmup 1 --no-c5 --no-parity --no-cpu-baseline --no-device-animation > $OUT/r02b_ncu_skin_$v.log 2>&1
done
echo "[v2] ncu of the C4 main level, variant 3 and 0"
for v in 3 0; do
  FYX_CULL_VARIANT=$v timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_update_level --launch-skip 3 -c 1 -o $OUT/r02b_full_c4_update_v$v \
      python bench.py --workload C4 --steps 2 --warmup 1 --no-c5 --no-parity --no-cpu-baseline --no-device-animation > $OUT/r02b_ncu_v$v.log 2>&1
done
ls -la $OUT/r02b_*.ncu-rep
echo "[v2] done"
