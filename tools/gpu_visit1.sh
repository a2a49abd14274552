#!/usr/bin/env bash
# Round-2 visit 1 (1 GPU): GPU tests incl. the full-size C3/C4 checks, the default bench (C4 + parity + strong C5 at N=1),
# compute-sanitizer memcheck / racecheck over the small parity tests.
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "[v1] tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $OUT/r02a_tests.log; tail -4 $OUT/r02a_tests.log
echo "[v1] bench default"; timeout 900 python bench.py --verbose > $OUT/r02a_bench_C4.json 2> $OUT/r02a_bench_C4.err; tail -5 $OUT/r02a_bench_C4.err; head -c 1500 $OUT/r02a_bench_C4.json; echo
SMALL='test_update_matches_oracle_on_generated_scenes or test_incremental_update or test_cull_six_cube_faces or test_palette_and_skinning or test_render_prep_one_call or test_pipelined_frames or test_skinned_mesh_box_quirk'
echo "[v1] memcheck"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "$SMALL" > $OUT/r02a_memcheck.log 2>&1; echo "rc=$?" >> $OUT/r02a_memcheck.log; tail -6 $OUT/r02a_memcheck.log
echo "[v1] racecheck"; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "$SMALL" > $OUT/r02a_racecheck.log 2>&1; echo "rc=$?" >> $OUT/r02a_racecheck.log; tail -6 $OUT/r02a_racecheck.log
echo "[v1] done"
