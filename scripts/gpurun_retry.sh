#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3: nothing charged).  usage: gpurun_retry.sh [gpurun options] -- command
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry $i] busy, sleeping 45 s"
  sleep 45
done
exit 3
