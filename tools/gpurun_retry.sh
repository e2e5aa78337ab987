#!/bin/bash
# usage: [GPUS=N] tools/gpurun_retry.sh <timeout> <command...>: retries while the pod answers busy / transient (exit code 3)
T=$1; shift
G=""; [ -n "$GPUS" ] && G="--gpus $GPUS"
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun $G --timeout $T -- "$@" > /tmp/gpurun_last.log 2>&1; rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" /tmp/gpurun_last.log; then break; fi
  sleep 60
done
tail -n ${TAILN:-60} /tmp/gpurun_last.log
exit $rc
