for args in "hot_threshold=128 pace_trail=2 chunk_events=1536" "hot_threshold=256 pace_trail=2 chunk_events=1280" "hot_threshold=256 pace_trail=2 chunk_events=1024" "hot_threshold=256 pace_trail=2 chunk_events=1536 ckpt_period=20" "hot_threshold=256 pace_trail=2 chunk_events=1536 bucket_width=0.03125 window_span=4" "hot_threshold=256 pace_trail=2 chunk_events=1536 substeps=12"; do
  echo "=== $args"
  timeout 300 python tools/gpu_probe.py 1048576 100 0 $args 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); s=d['stats']
        print('ev/s=%.4g device_s=%.3f windows=%d rounds=%d rollbacks=%d silent=%d executed=%d launches=%d status=%s' % (d['events_per_s_device'], s['device_seconds'], s['windows'], s['relax_rounds'], s['rollbacks'], s['silent_events'], s['executed_events'], s['kernel_launches'], s['status']))
"
done
