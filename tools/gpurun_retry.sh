#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3, nothing charged):  tools/gpurun_retry.sh [gpurun args] -- 'command'
for attempt in $(seq 1 30); do
    /usr/local/graft/bin/gpurun "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    echo "[retry] attempt $attempt answered busy; sleeping 90 s" >&2
    sleep 90
done
exit 3
