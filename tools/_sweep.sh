set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for args in "" "chunk_events=2048" "window_span=2" "pace_trail=2"; do
  echo "=== $args"
  timeout 300 python tools/gpu_probe.py 1048576 100 1 $args 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); s=d['stats']
        print('ev/s=%.4g device_s=%.3f windows=%d rounds=%d rollbacks=%d silent=%d executed=%d launches=%d status=%s' % (d['events_per_s_device'], s['device_seconds'], s['windows'], s['relax_rounds'], s['rollbacks'], s['silent_events'], s['executed_events'], s['kernel_launches'], s['status_names']))
        print(' '.join('%s=%.0f/%d' % (k, v[0], v[1]) for k, v in list(d['kernel_ms'].items())[:9]))
"
done
RS_DEBUG=2 timeout 300 python tools/gpu_probe.py 1048576 100 0 2> gpurun_out/r2_winlog_b.txt | tail -1 | cut -c1-300
RS_DISTINCT_SEEDS=1 timeout 300 python tools/gpu_probe.py 1048576 100 1 2>/dev/null | tail -1 | cut -c1-1500
