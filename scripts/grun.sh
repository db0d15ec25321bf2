#!/bin/bash
# scripts/grun.sh LOG TIMEOUT 'command' [gpus] -- gpurun with retries while the pod answers "busy" (exit 3)
LOG=$1; TMO=$2; CMD=$3; GP=${4:-1}
for try in $(seq 1 30); do
  if [ "$GP" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $TMO -- "$CMD" > $LOG 2>&1; else /usr/local/graft/bin/gpurun --gpus $GP --timeout $TMO -- "$CMD" > $LOG 2>&1; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
