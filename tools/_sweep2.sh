#!/bin/bash
# usage: tools/_sweep2.sh "ENV=.. args" "..." ; each entry: optional VAR=val env assignments (uppercase) followed by cfg overrides
for entry in "$@"; do
  envs=""; args=""
  for tok in $entry; do
    case "$tok" in
      [A-Z]*=*) envs="$envs $tok";;
      *) args="$args $tok";;
    esac
  done
  echo "=== $entry"
  env $envs timeout 300 python tools/gpu_probe.py 1048576 100 0 $args 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); s=d['stats']
        print('ev/s=%.4g device_s=%.3f windows=%d rounds=%d rollbacks=%d silent=%d executed=%d launches=%d status=%s' % (d['events_per_s_device'], s['device_seconds'], s['windows'], s['relax_rounds'], s['rollbacks'], s['silent_events'], s['executed_events'], s['kernel_launches'], s['status']))
"
done
