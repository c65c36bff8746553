#!/bin/bash
# usage: retry.sh <logfile> <gpurun args...>   retries ONLY while the pod answers busy (exit 3 with fault "busy");
# a lost box (fault "lease") is a strike and is never retried automatically
log=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  fault=$(python -c "import json;print(json.load(open('gpurun_out/.last_call.json')).get('fault',''))" 2>/dev/null)
  if [ "$fault" != "busy" ]; then exit $rc; fi
  sleep 120
done
exit 3
