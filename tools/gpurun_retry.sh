#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> '<command>'   — retries while the pod answers busy/transient (nothing charged)
T=$1; shift
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  echo "$out" | tail -80
  if echo "$out" | grep -q "status=transient\|rc=3\|nothing was charged"; then
    echo "[retry $i] busy; sleeping 150 s"; sleep 150; continue
  fi
  break
done
