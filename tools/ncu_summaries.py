"""Turn the raw ncu outputs of a GPU session into the small text summaries committed under profiles/.
  python tools/ncu_summaries.py launches <launches.csv> <out.txt> "<title>"
  python tools/ncu_summaries.py full <report.ncu-rep> <out.txt> "<title>"      (needs ncu on PATH)
"""
import collections
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]


def launches(path, out, title):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    tot = collections.Counter()
    cnt = collections.Counter()
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[iu], 1.0)
        name = r[ik].split("(")[0]
        tot[name] += v
        cnt[name] += 1
    allns = sum(tot.values())
    with open(out, "w") as f:
        f.write("# %s\n# ncu --metrics gpu__time_duration.sum --clock-control none ; per-launch times are cold-cache and serialised:\n"
                "# compare SHARES with the live CUDA-event numbers in r01_bench_1gpu.json (kernel_ms), not absolutes\n"
                "# kernel, launches, total (ns), share\n" % title)
        for k, v in tot.most_common():
            f.write("%-22s %6d %14.1f %6.2f%%\n" % (k, cnt[k], v, 100.0 * v / allns))


def full(rep, out, title):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        e = {"Kernel Name": d["Kernel Name"]}
        for k in KEYS:
            if k in d:
                e[k] = d[k] + " " + units[hdr.index(k)]
        res.append(e)
    with open(out, "w") as f:
        f.write("# %s\n" % title)
        f.write(json.dumps(res, indent=1) + "\n")
    return res


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4])
