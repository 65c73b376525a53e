#!/usr/bin/env python3
"""Per-event instruction budget of a kernel from an `ncu --set full` report: every SASS instruction's executed
count (ncu --page source --csv) joined with the cubin's line table (nvdisasm --print-line-info-inline), aggregated
by the source line of the OUTERMOST frame inside a given function range and by innermost line.

usage: python tools/ncu_event_budget.py report.ncu-rep lib.so kernel_substring events [lo hi]
  events = number of events the launch executed (divides the counts); lo..hi = line range of rs_engine.cuh to list"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def load(rep, lib, kern):
    work = tempfile.mkdtemp(prefix="ncu-bud-")
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=work, capture_output=True)
    cubin = [f for f in os.listdir(work) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(work, cubin)], capture_output=True, text=True).stdout.split("\n")
    insts, inside, pending, cur = [], False, [], [("?", 0)]
    for ln in sass:
        if ln.startswith(".text."):
            inside = kern in ln
            pending = []
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            pending.append(m)
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if m:
            if pending:
                st = [(os.path.basename(pending[0].group(1)), int(pending[0].group(2)))]
                for p in pending:
                    if p.group(3):
                        st.append((os.path.basename(p.group(3)), int(p.group(4))))
                cur = st
                pending = []
            insts.append((int(m.group(1), 16), cur, m.group(2).strip()))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.split("\n")))
    hdr = rows[1]
    ia, ie, ism = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall = {h: i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
    recs = [r for r in rows[2:] if len(r) == len(hdr)]
    base = int(recs[0][ia], 16)
    by_off = {off: (st, txt) for off, st, txt in insts}
    res = []
    for r in recs:
        off = int(r[ia], 16) - base
        st, txt = by_off.get(off, ([("?", 0)], "?"))
        res.append(dict(off=off, stack=st, sass=txt, execs=int(r[ie]), samples=int(r[ism]), stalls={k: int(r[i]) for k, i in stall.items() if int(r[i])}))
    return res


def main():
    rep, lib, kern, events = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    res = load(rep, lib, kern)
    tot = sum(x["execs"] for x in res)
    smp = sum(x["samples"] for x in res)
    print("instructions executed %d = %.1f per event; samples %d" % (tot, tot / events, smp))
    inner = collections.Counter(); inner_s = collections.Counter()
    outer = collections.Counter(); outer_s = collections.Counter()
    for x in res:
        inner[x["stack"][0]] += x["execs"]; inner_s[x["stack"][0]] += x["samples"]
        # outermost frame that lies in rs_engine.cuh inside the visit function (the last but one frame is the kernel's call site)
        fr = [f for f in x["stack"] if f[0] == "rs_engine.cuh"]
        key = fr[-2] if len(fr) >= 2 else (fr[-1] if fr else x["stack"][-1])
        outer[key] += x["execs"]; outer_s[key] += x["samples"]
    print("--- by line of the visit function (outer frame): inst/event, share of samples")
    for k, v in sorted(outer.items(), key=lambda kv: -outer_s[kv[0]])[:70]:
        print("%-22s %7.1f inst/ev  %5.1f%% samples" % ("%s:%d" % k, v / events, 100.0 * outer_s[k] / smp))
    print("--- by innermost line")
    for k, v in sorted(inner.items(), key=lambda kv: -inner_s[kv[0]])[:70]:
        print("%-22s %7.1f inst/ev  %5.1f%% samples" % ("%s:%d" % k, v / events, 100.0 * inner_s[k] / smp))


if __name__ == "__main__":
    main()
