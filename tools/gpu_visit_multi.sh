#!/usr/bin/env bash
# Multi-GPU visit (run with gpurun --gpus N): exchange variants through the parity test, then bench.py under torchrun with the
# default exchange (peer stores + host segment) and with the round-1 form (NCCL, private host copies) for comparison.
#   usage: bash tools/gpu_visit_multi.sh <N> <tag> [full|compare|quick|benchonly]
set -u
N=${1:-2}
TAG=${2:-r02m}
MODE=${3:-full}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L | head -8
if [ "$MODE" != "benchonly" ]; then
echo "[multi] parity test of every exchange variant (2 ranks)"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -12 > $OUT/${TAG}_tests.log; tail -12 $OUT/${TAG}_tests.log
fi
run() { # name, extra env, extra args
  echo "[multi] bench N=$N $1"
  env $2 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 --verbose $3 \
      > $OUT/${TAG}_bench_$1.json 2> $OUT/${TAG}_bench_$1.err
  grep -E "parity|scene on device" $OUT/${TAG}_bench_$1.err | cut -c1-400
  python - "$OUT/${TAG}_bench_$1.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    st = {k: (round(v["ms"], 4), round(v["frac"], 3)) for k, v in d["roofline"]["stages"].items()}
    print("   C4 weak: ms/frame", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), "sync", d["e2e"]["ms_per_step_synchronous"], st, "parity", d["parity"]["ok"] if d.get("parity") else None)
    print("   exchange:", d["config"]["exchange"])
    c5 = d["modes"].get("strong_C5")
    if c5:
        print("   C5 strong:", {k: c5.get(k) for k in ("ms_per_step", "e2e_ms_per_step", "value", "e2e_value")}, "parity", (c5.get("parity") or {}).get("ok"), c5.get("error"))
except Exception as ex:
    print("   (no JSON line)", ex)
PY
}
run default "FYX_DUMMY=1" ""
if [ "$MODE" = "compare" ]; then
  run nccl_private "FYX_EXCHANGE=nccl FYX_HOSTSEG=0" "--no-c5"
fi
if [ "$MODE" = "full" ]; then
  run nccl_private "FYX_EXCHANGE=nccl FYX_HOSTSEG=0" "--no-c5"
  run nccl_seg "FYX_EXCHANGE=nccl" "--no-c5"
  run peer_private "FYX_HOSTSEG=0" "--no-c5"
fi
echo "[multi] done"
