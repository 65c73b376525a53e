"""Aggregate an `ncu --page source --csv` dump: samples per source line and the top SASS lines.
usage: python tools/ncu_src_agg.py src.csv [topN] [--sass]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 40
hdr = None; cur = None
agg = collections.Counter(); src = {}; exe = collections.Counter()
sass = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1]; continue
    if r and r[0] in ("Line No", "Address", "#"): hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try: n = int(d["# Samples"]); b = int(d.get("stall_barrier") or 0)
        except Exception: continue
        if "Line No" in d:
            key = (cur.split('/')[-1] if cur else '?', d["Line No"])
            agg[key] += n - b; src[key] = r[1][:120]
            try: exe[key] += int(d["Instructions Executed"])
            except Exception: pass
        else:
            sass.append((n - b, d))
tot = sum(v for k, v in agg.items() if k[1] != '')
print("total non-barrier samples (lines):", tot)
for k, v in agg.most_common(top):
    print("%6d %5.1f%% x%-8d %s:%s  %s" % (v, 100.0 * v / max(tot, 1), exe[k], k[0], k[1], src[k]))
