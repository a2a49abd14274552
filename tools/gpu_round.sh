#!/usr/bin/env bash
# One GPU-box visit that produces everything a round needs (box acquisition is the expensive part of a gpurun call):
#   tests -> bench (C4 default, C2, target, C3) -> ncu launch list of the default bench command -> ncu --set full of the
#   hot kernels -> N2 / N3 tools.  Everything lands in gpurun_out/<tag>_*; summaries: tools/make_profiles.py.
#   usage (from the repo root, on the GPU box):  bash tools/gpu_round.sh r02 [quick]
# "quick" skips the ncu captures.  Every step has its own timeout so that a hang cannot eat the budget.
set -u
TAG=${1:-rXX}
MODE=${2:-full}
OUT=gpurun_out
mkdir -p $OUT
log() { echo "[gpu_round] $*"; }

log "tests"
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $OUT/${TAG}_tests.log
tail -3 $OUT/${TAG}_tests.log

for w in C4 C2 target C3; do
    extra="--no-cpu-baseline --no-device-animation"
    [ "$w" = "C4" ] && extra=""   # the default command, complete
    log "bench $w"
    timeout 600 python bench.py --workload $w $extra > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err
    python - "$OUT/${TAG}_bench_$w.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    st = {k: (round(v["ms"], 4), round(v["frac"], 3)) for k, v in d["roofline"]["stages"].items()}
    print("   ms/frame", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), st)
except Exception as ex:
    print("   (no JSON line)", ex)
PY
done

log "reference arm"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err

log "N3 / N2 tools"
timeout 200 python tools/drawprep_bench.py > $OUT/${TAG}_drawprep.json 2> $OUT/${TAG}_drawprep.err
cat $OUT/${TAG}_drawprep.json

if [ "$MODE" != "quick" ]; then
    K='regex:k_update_level|k_fold|k_palette|k_skin|k_snapshot|k_cull|k_scatter_trs|k_anim|k_inst|k_lod'
    log "ncu launch list of the default bench command"
    timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --csv --log-file $OUT/${TAG}_launches_C4.csv \
        python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-device-animation > $OUT/${TAG}_launches.log 2>&1
    log "ncu --set full: main level of the fused update+cull (C4), skin + palette + fold (target)"
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_update_level --launch-skip 3 -c 1 -o $OUT/${TAG}_full_c4_update \
        python bench.py --workload C4 --steps 2 --warmup 1 --no-cpu-baseline --no-device-animation > $OUT/${TAG}_full_a.log 2>&1
    timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:k_skin|k_palette|k_fold' --launch-skip 6 -c 3 -o $OUT/${TAG}_full_target_skin \
        python bench.py --workload target --steps 2 --warmup 1 --no-cpu-baseline --no-device-animation > $OUT/${TAG}_full_b.log 2>&1
    log "ncu --set full: N2 (C3 device animation), N3 (C2 draw-prep)"
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_anim_ --launch-skip 9 -c 3 -o $OUT/${TAG}_full_anim \
        python bench.py --workload C3 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_full_c.log 2>&1
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_inst_ --launch-skip 6 -c 3 -o $OUT/${TAG}_full_inst \
        python tools/drawprep_bench.py --steps 3 > $OUT/${TAG}_full_d.log 2>&1
    ls -la $OUT/${TAG}_*.ncu-rep 2>/dev/null
fi
log "done"
