#!/usr/bin/env bash
# Round-2 visit 15 (1 GPU): the "static + skeletons" frame (FYX_UPDATE_INCREMENTAL) against the all-dirty frame, stage by stage,
# for cull variants 20 (non-UA path = variant 0), 21 (-> pre-reject), 52 (-> warp-convergent predicate).
set -u
OUT=gpurun_out
mkdir -p $OUT
for v in 20 21 52 20 21 52; do
  echo "[v15] variant $v"
  FYX_CULL_VARIANT=$v timeout 300 python tools/inc_stages.py 50000 2> $OUT/r02u_v$v.err | tee -a $OUT/r02u_inc_stages.jsonl | cut -c1-900
done
echo "[v15] done"
