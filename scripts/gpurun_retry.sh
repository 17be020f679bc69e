#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> '<command>'   -- retries while the pod has no free GPU slot
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  if ! python3 -c "import json,sys; d=json.load(open('gpurun_out/.last_call.json')); sys.exit(0 if d.get('status')=='transient' else 1)" 2>/dev/null; then exit $rc; fi
  sleep 60
done
exit 3
