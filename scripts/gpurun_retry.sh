#!/bin/bash
# usage: scripts/gpurun_retry.sh <log> <timeout-s> <command...>   -- retries while the pod answers busy / transient
LOG=$1; shift; TMO=$1; shift
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $TMO -- "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" $LOG || [ $rc -eq 3 ]; then sleep 60; continue; fi
  exit $rc
done
exit 3
