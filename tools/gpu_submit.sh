#!/bin/bash
# Submits one gpurun call and retries while the pod answers "busy" (exit code 3: nothing charged).
# Usage: tools/gpu_submit.sh TAG TIMEOUT [--gpus N] -- steps...   (steps as in tools/gpu_run.sh)
TAG=$1; TMO=$2; shift 2
GP=""
if [ "$1" == "--gpus" ]; then GP="--gpus $2"; shift 2; fi
[ "$1" == "--" ] && shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun $GP --timeout $TMO -- "bash tools/gpu_run.sh $TAG $*" > gpurun_out/${TAG}_call.log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" gpurun_out/${TAG}_call.log; then break; fi
  sleep 45
done
echo "exit $rc" >> gpurun_out/${TAG}_call.log
