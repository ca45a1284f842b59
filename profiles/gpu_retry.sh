#!/bin/bash
# gpu_retry.sh <log> <gpurun args...> : retry a gpurun call while the pod answers "busy" (exit code 3), up to 12 times
log=$1; shift
for attempt in $(seq 1 12); do
    /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
    rc=$?
    if [ $rc -ne 3 ]; then echo "gpurun rc=$rc attempt=$attempt" >> "$log"; exit $rc; fi
    sleep 150
done
echo "gpurun still busy after 12 attempts" >> "$log"; exit 3
