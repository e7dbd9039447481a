#!/bin/bash
# Local helper: run one gpurun call, retrying while the pod answers "transient" (no slot; nothing charged).
# usage: tools/gpurun_retry.sh <log> <gpurun args...>
LOG=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  if ! grep -q "status=transient" "$LOG"; then break; fi
  sleep 90
done
tail -4 "$LOG" | cut -c1-400
