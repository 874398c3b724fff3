#!/bin/bash
# usage: tools/gpurun_retry.sh <gpurun args...>   - retries while the pod answers "busy" (exit code 3)
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] attempt $attempt answered busy; sleeping 120 s"
  sleep 120
done
exit 3
