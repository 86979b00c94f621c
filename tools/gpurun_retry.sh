#!/bin/bash
# usage: tools/gpurun_retry.sh <outfile> <timeout> [--gpus N] -- <command>
# retries while the pod answers "busy" (exit 3 / status=transient); nothing is charged for those
out=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to "$@" > $out 2>&1; rc=$?
  if grep -q "status=transient" $out || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
echo "gpurun_retry done rc=$rc attempts=$i" >> $out
