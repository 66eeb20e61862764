#!/bin/bash
# usage: tools/gpucall.sh <name> <timeout> [--gpus N] -- <command>   (retries while the pod has no free slot)
name=$1; to=$2; shift 2
gp=""
if [ "$1" = "--gpus" ]; then gp="--gpus $2"; shift 2; fi
shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to $gp -- "$@" > gpurun_out/$name.out 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" gpurun_out/$name.out; then exit $rc; fi
  sleep 60
done
exit 3
