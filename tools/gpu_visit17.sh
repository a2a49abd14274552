#!/usr/bin/env bash
# Round-2 visit 17 (1 GPU): what the driver runs at round end, on the final tree: GPU tests, smoke(), the default bench command and
# the reference arm.
set -u
OUT=gpurun_out
mkdir -p $OUT
if [ "${1:-all}" = "all" ]; then
echo "[v17] pytest -m gpu"; timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
echo "[v17] smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
fi
echo "[v17] default bench"; T0=$SECONDS; timeout 900 python bench.py > $OUT/r02w_bench_default.json 2> $OUT/r02w_bench_default.err; echo "   wall $((SECONDS-T0)) s"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02w_bench_default.json"))
print("   ", d["metric"], d["value"], d["unit"], "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 3), "launches", d["gpu_launches"], "parity", d["parity"]["ok"], "clocks", d["clocks"])
PY
echo "[v17] reference arm"; T0=$SECONDS; timeout 900 python bench.py --impl reference > $OUT/r02w_bench_reference.json 2> $OUT/r02w_bench_reference.err; echo "   wall $((SECONDS-T0)) s"; cut -c1-400 $OUT/r02w_bench_reference.json
echo "[v17] done"
