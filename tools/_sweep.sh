for lib in libphold_r4 libphold_r3 libphold_r6; do
  echo "=== $lib"
  RS_LIB=root-sim_b200/_built/$lib.so timeout 300 python tools/gpu_probe.py 1048576 100 1 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); s=d['stats']
        print('ev/s=%.4g device_s=%.3f windows=%d rounds=%d rollbacks=%d silent=%d executed=%d launches=%d status=%s ra=%d cuts=%d' % (d['events_per_s_device'], s['device_seconds'], s['windows'], s['relax_rounds'], s['rollbacks'], s['silent_events'], s['executed_events'], s['kernel_launches'], s['status'], s['run_ahead_lps'], s['window_cuts']))
        print(' '.join('%s=%.0f/%d' % (k, v[0], v[1]) for k, v in list(d['kernel_ms'].items())[:9]))
"
  RS_DISTINCT_SEEDS=1 RS_LIB=root-sim_b200/_built/$lib.so timeout 300 python tools/gpu_probe.py 1048576 100 1 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); s=d['stats']
        print('DISTINCT ev/s=%.4g device_s=%.3f rounds=%d' % (d['events_per_s_device'], s['device_seconds'], s['relax_rounds']), ' '.join('%s=%.0f/%d' % (k, v[0], v[1]) for k, v in list(d['kernel_ms'].items())[:6]))
"
done
