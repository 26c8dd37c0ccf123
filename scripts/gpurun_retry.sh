#!/usr/bin/env bash
# gpurun with retries while the pod has no free GPU slot (exit code 3 = nothing charged): scripts/gpurun_retry.sh <gpurun args...>
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpurun_retry] busy, attempt $i; sleeping 90 s" >&2
  sleep 90
done
exit 3
