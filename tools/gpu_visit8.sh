#!/usr/bin/env bash
# Round-2 visit 8 (1 GPU): k_skin2 (two vertices per thread) against k_skin on one box.
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "[v8] skin variants tests"; timeout 900 python -m pytest tests/test_gpu_variants.py -m gpu -q -k "tma or pair" 2>&1 | tail -4
for v in ldg pair4 pair5 pair6 ldg pair5; do
  for w in C4 target; do
    echo "[v8] skin variant $v workload $w"
    FYX_SKIN_VARIANT=$v timeout 300 python bench.py --workload $w --no-c5 --no-parity --no-cpu-baseline --no-device-animation --steps 20 > $OUT/r02j_${v}_$w.json 2> $OUT/r02j_${v}_$w.err
    python - "$OUT/r02j_${v}_$w.json" <<'PY'
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
FYX_SKIN_VARIANT=pair5 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_skin --launch-skip 2 -c 1 -o $OUT/r02j_full_skin_pair5 \
    python bench.py --workload target --steps 2 --warmup 1 --no-c5 --no-parity --no-cpu-baseline --no-device-animation > $OUT/r02j_ncu_pair5.log 2>&1
ls -la $OUT/r02j_*.ncu-rep
echo "[v8] done"
