#!/usr/bin/env bash
# Round-2 visit 11 (1 GPU): compute-sanitizer memcheck over the draw-prep / animation / fuzz tests; initcheck over the small parity tests.
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "[v11] memcheck: drawprep + anim + fuzz(seed 1)"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_drawprep.py tests/test_gpu_anim.py tests/test_gpu_fuzz.py -q -x -k "not seed2 and not 2] and not 3] and not 4]" > $OUT/r02n_memcheck2.log 2>&1; echo "rc=$?" >> $OUT/r02n_memcheck2.log; tail -5 $OUT/r02n_memcheck2.log
SMALL='test_update_matches_oracle_on_generated_scenes or test_incremental_update or test_cull_six_cube_faces or test_palette_and_skinning or test_render_prep_one_call or test_pipelined_frames'
echo "[v11] initcheck (informational)"
timeout 900 compute-sanitizer --tool initcheck --show-backtrace no --print-limit 3000 --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "$SMALL" > $OUT/r02n_initcheck.log 2>&1; echo "rc=$?" >> $OUT/r02n_initcheck.log; tail -8 $OUT/r02n_initcheck.log
echo "[v11] done"
