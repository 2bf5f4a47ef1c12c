#!/bin/bash
# usage: tools/gpurun_retry.sh [--gpus N] <timeout_s> '<command>'  -- retries while the pod answers busy (exit code 3)
gp=""
if [ "$1" = "--gpus" ]; then gp="--gpus $2"; shift 2; fi
to=$1; shift
for i in $(seq 1 40); do
  # stdin from /dev/null: a tool that waits on stdin must fail, not burn the GPU budget until the limit
  /usr/local/graft/bin/gpurun $gp --timeout "$to" -- "exec < /dev/null; $*"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
