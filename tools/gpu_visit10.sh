#!/usr/bin/env bash
# Round-2 visit 10 (1 GPU): the randomised differential test, the C++ host mirror test, k_skin alone (sustained vs cold).
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "[v10] fuzz + mirror tests"; timeout 1800 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -15
echo "[v10] k_skin alone"; timeout 600 python tools/skin_alone.py 50000 > $OUT/r02m_skin_alone.json 2> $OUT/r02m_skin_alone.err; cat $OUT/r02m_skin_alone.json; tail -2 $OUT/r02m_skin_alone.err
echo "[v10] done"
