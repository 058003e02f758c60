#!/bin/bash
# usage: scripts/gpurun_retry.sh LOGFILE TIMEOUT -- command...   (retries while the pod answers "busy")
LOG=$1; TO=$2; shift 3
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1
  if grep -q "status=transient" $LOG; then sleep 90; else break; fi
done
