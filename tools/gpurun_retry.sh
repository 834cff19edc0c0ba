#!/bin/bash
# gpurun with retries on "no box or slot free right now" (exit code 3: nothing is charged).  Usage: tools/gpurun_retry.sh <gpurun args...>
# The repository is snapshotted at every attempt: do not edit sources the queued command uses while this is waiting.
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpurun_retry] attempt $attempt: busy, retrying in 150 s" >&2
  sleep 150
done
exit 3
