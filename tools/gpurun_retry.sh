#!/bin/bash
# development aid: retry a gpurun call while the pod answers "busy" (exit code 3 / transient)
# usage: tools/gpurun_retry.sh LOGFILE [gpurun options] -- 'command'
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if ! grep -q "status=transient\|no box or slot" "$log"; then exit $rc; fi
  sleep 120
done
exit 3
