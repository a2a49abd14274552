#!/usr/bin/env bash
# Round-2 visit 16 (1 GPU): incremental path with the stored box / static columns requested up front — parity tests, then the
# stage tool (compare with visit r02u: update+cull 0.348 ms incremental, 0.415 all-dirty).
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "[v16] tests"; timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_drawprep.py tests/test_gpu_anim.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -4
for v in 20 20; do
  echo "[v16] variant $v"
  FYX_CULL_VARIANT=$v timeout 300 python tools/inc_stages.py 50000 2> $OUT/r02v_v$v.err | tee -a $OUT/r02v_inc_stages.jsonl | cut -c1-900
done
echo "[v16] target + C2 incremental-free sanity (default bench, quick)"
timeout 300 python bench.py --workload target --no-c5 --no-cpu-baseline --no-device-animation --steps 20 > $OUT/r02v_target.json 2> $OUT/r02v_target.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02v_target.json"))
print("   target ms/frame", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), "inc e2e", d["modes"]["static_plus_skeletons_e2e_ms_per_step"], "parity", d["parity"]["ok"])
PY
echo "[v16] done"
