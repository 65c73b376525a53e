#!/usr/bin/env python3
"""Map the per-instruction samples of an `ncu --page source --csv` (SASS view) dump to source lines, using the
line table of the cubin (`nvdisasm --print-line-info-inline`): the sources need not be present where ncu ran.

usage: python tools/ncu_sass_lines.py report.ncu-rep lib.so kernel_substring [topN]
Innermost line (where the instruction comes from) and outermost line (where it was inlined into the kernel) are
both aggregated."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    work = tempfile.mkdtemp(prefix="ncu-sass-")
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=work, capture_output=True)
    cubin = [f for f in os.listdir(work) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(work, cubin)], capture_output=True, text=True).stdout.split("\n")
    # instruction list of the kernel: (offset, innermost (file, line), outermost (file, line))
    insts = []
    inside = False
    pending = []
    cur = (("?", 0), ("?", 0))
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
                inner = (os.path.basename(pending[0].group(1)), int(pending[0].group(2)))
                last = pending[-1]
                outer = (os.path.basename(last.group(3) or last.group(1)), int(last.group(4) or last.group(2)))
                cur = (inner, outer)
                pending = []
            insts.append((int(m.group(1), 16), cur, m.group(2).strip()))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.split("\n")))
    hdr = None
    samples = []
    for r in rows:
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            samples.append(d)
    # align by address: ncu gives absolute addresses, the listing offsets from the start of the kernel
    base = int(samples[0]["Address"], 16)
    by_off = {off: (lines, text) for off, lines, text in insts}
    aligned = []
    for d in samples:
        off = int(d["Address"], 16) - base
        lines, text = by_off.get(off, ((("?", 0), ("?", 0)), "?"))
        aligned.append((off, lines, text))
    insts = aligned
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    inner_agg, outer_agg = collections.Counter(), collections.Counter()
    inner_exec = collections.Counter()
    stall_by_line = collections.defaultdict(collections.Counter)
    tot = 0
    for d, (off, (inner, outer), text) in zip(samples, insts):
        n = int(d["# Samples"]) - int(d.get("stall_barrier") or 0)
        tot += n
        inner_agg[inner] += n
        outer_agg[outer] += n
        inner_exec[inner] += int(d["Instructions Executed"] or 0)
        for s in stall_cols:
            if s != "stall_barrier":
                stall_by_line[inner][s] += int(d[s] or 0)
    srcs = {}

    def src(fl):
        f, l = fl
        if f not in srcs:
            for cand in (os.path.join("root-sim_b200", "csrc", f), os.path.join("root-sim_b200", "include", f)):
                if os.path.exists(cand):
                    srcs[f] = open(cand).read().split("\n")
                    break
            else:
                srcs[f] = []
        return srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""

    print("non-barrier samples:", tot)
    print("--- by innermost source line")
    for k, v in inner_agg.most_common(top):
        st = ", ".join("%s %d" % (s.replace("stall_", ""), c) for s, c in stall_by_line[k].most_common(2))
        print("%6d %5.1f%% x%-9d %s:%d  %s   [%s]" % (v, 100.0 * v / max(tot, 1), inner_exec[k], k[0], k[1], src(k), st))
    print("--- by line of the kernel / visit function the code was inlined into")
    for k, v in outer_agg.most_common(top):
        print("%6d %5.1f%% %s:%d  %s" % (v, 100.0 * v / max(tot, 1), k[0], k[1], src(k)))


if __name__ == "__main__":
    main()
