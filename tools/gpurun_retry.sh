#!/bin/bash
# tools/gpurun_retry.sh [gpurun options] -- 'command' : gpurun with retries while the pod answers "busy" (exit 3)
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
