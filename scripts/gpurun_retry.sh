#!/bin/bash
# usage: scripts/gpurun_retry.sh [gpurun args...] -- 'command'   : retries while the pod answers busy / transient (nothing charged)
for i in 1 2 3 4 5 6 7 8; do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | tail -40
  if echo "$out" | grep -q "status=transient\|status=busy\|exit code 3"; then sleep 120; continue; fi
  break
done
