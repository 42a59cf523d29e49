#!/bin/bash
LOG=$1; shift; G=$1; shift; T=$1; shift
for n in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" > $LOG 2>&1
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 150
done
exit 3
