#!/usr/bin/env bash
# gpurun with retries while the pod answers "transient" (busy / backing off).  usage: tools/gpurun_retry.sh <log> <gpurun args...>
LOG=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  if grep -q "status=transient" "$LOG"; then
    w=$(grep -o "retry in [0-9]*s" "$LOG" | grep -o "[0-9]*" | head -1); w=${w:-200}
    echo "[retry $i] transient, sleeping $((w + 15)) s" >> "$LOG.retries"
    sleep $((w + 15))
  else
    break
  fi
done
tail -70 "$LOG"
