#!/bin/bash
# usage: tools/gpurun_retry.sh <gpurun args...>   -- retries while the pod answers busy (exit 3 / transient)
for attempt in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1); rc=$?
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient" || [ $rc -eq 3 ]; then
    echo "[retry] attempt $attempt busy; sleeping 120 s"; sleep 120; continue
  fi
  exit $rc
done
exit 3
