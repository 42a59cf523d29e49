#!/bin/bash
# tools/gpurun_retry.sh <log> <timeout> <command...>: keep asking for a B200 box until one is free (exit 3 = none right now)
LOG=$1; shift; T=$1; shift
for n in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > $LOG 2>&1
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 150
done
exit 3
