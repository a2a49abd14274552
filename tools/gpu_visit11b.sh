#!/usr/bin/env bash
# Round-2 visit 11b (1 GPU): initcheck over the small parity tests, no backtraces.
set -u
OUT=gpurun_out
mkdir -p $OUT
SMALL='test_update_matches_oracle_on_generated_scenes or test_incremental_update or test_cull_six_cube_faces or test_palette_and_skinning or test_render_prep_one_call or test_pipelined_frames'
echo "[v11b] initcheck (informational)"
timeout 900 compute-sanitizer --tool initcheck --show-backtrace no --print-limit 3000 --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "$SMALL" > $OUT/r02o_initcheck.log 2>&1; echo "rc=$?" >> $OUT/r02o_initcheck.log; tail -8 $OUT/r02o_initcheck.log
echo "[v11b] done"
