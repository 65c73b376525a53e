set -x
for args in "" "hot_threshold=2048" "hot_threshold=2048 chunk_events=2048" "window_span=2 hot_threshold=1024"; do
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
timeout 600 ncu --set full --import-source on -k regex:k_run -s 850 -c 1 -f -o gpurun_out/r2_krun_late python tools/gpu_probe.py 1048576 100 0 > gpurun_out/ncu_late.log 2>&1; tail -2 gpurun_out/ncu_late.log
timeout 600 ncu --set full --import-source on -k regex:k_run -s 120 -c 1 -f -o gpurun_out/r2_krun_early python tools/gpu_probe.py 1048576 100 0 > gpurun_out/ncu_early.log 2>&1; tail -2 gpurun_out/ncu_early.log
RS_DEBUG=2 timeout 300 python tools/gpu_probe.py 1048576 100 0 2> gpurun_out/r2_winlog_c.txt | tail -1 | cut -c1-300
