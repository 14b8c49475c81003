#!/bin/bash
# Dev: call gpurun until the pod has a slot (exit code 3 = busy, nothing charged).  usage: tools/gpurun_retry.sh <gpurun args...>
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] pod busy, attempt $i; sleeping 45 s"
  sleep 45
done
exit 3
