#!/bin/bash
# retry a gpurun call while the pod answers "busy / transient" (nothing is charged for those); usage: gpurun_retry.sh <timeout> <command...>
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout $T -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient\|retry in a few minutes" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
cat /tmp/gpurun_last.log
exit $rc
