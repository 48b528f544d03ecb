#!/bin/bash
# usage: tools/gpu/retry.sh TIMEOUT 'command'   -- re-submit while the pod answers "transient" (nothing is charged for those)
T=$1; shift
cd /root/repo
# the snapshot must be coherent: rebuild the library (no-op when up to date) before every submission
python build.py > /tmp/retry_build.log 2>&1 || { tail -5 /tmp/retry_build.log; exit 1; }
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient"; then sleep 150; continue; fi
  break
done
