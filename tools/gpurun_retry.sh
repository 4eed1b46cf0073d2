#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun options] -- 'command'   (retries while the pod answers "busy", exit code 3)
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
