#!/bin/bash
# retry a gpurun call while the pod answers "busy" (status=transient: nothing charged); usage: tools/gpurun_retry.sh TRIES GPURUN-ARGS...
tries=$1; shift
for i in $(seq 1 $tries); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | tail -40
  if ! echo "$out" | grep -q "status=transient"; then exit 0; fi
  echo "--- attempt $i: busy, retrying in 200 s"
  sleep 200
done
exit 3
