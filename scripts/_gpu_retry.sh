#!/bin/bash
# usage: scripts/_gpu_retry.sh <logfile> <gpurun args...> : retries while the pool answers "busy" (exit code 3)
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
