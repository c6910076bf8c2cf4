#!/bin/bash
# usage: tools/gpurun_retry.sh LOGFILE TIMEOUT [--gpus N] -- 'command'    retries while the pod answers "busy" (exit 3)
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "rc=$rc" >> $log; exit $rc; fi
  sleep 60
done
echo "rc=3 gave up" >> $log
