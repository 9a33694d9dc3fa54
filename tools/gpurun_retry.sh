#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <script> <log>   -- retries while the pod answers busy (nothing charged)
for attempt in 1 2 3 4 5 6 7 8 9 10; do
  /usr/local/graft/bin/gpurun --timeout $1 -- "bash $2 > $3 2>&1" > /tmp/gpurun_last.out 2>&1
  if grep -q "status=ok\|status=fail\|status=timeout" /tmp/gpurun_last.out; then break; fi
  sleep 90
done
cat /tmp/gpurun_last.out | tail -4
