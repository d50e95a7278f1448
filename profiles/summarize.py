"""Turns gpurun_out/<launch list>.csv (+ an optional .ncu-rep) into the committed profiles/*.md summary.
usage: python profiles/summarize.py <launches.csv> [<report.ncu-rep>] > profiles/rNN_<what>.md"""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
    agg = collections.OrderedDict()
    for r in rows:
        k = r["Kernel Name"].split("(")[0].replace("void ", "")
        a = agg.setdefault(k, [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += float(r["Metric Value"])
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total ms | avg us | share | grid | block |")
    print("|---|---:|---:|---:|---:|---|---|")
    for k, (c, t, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.1f | %.1f%% | %s | %s |" % (k, c, t / 1e6, t / c / 1e3, 100 * t / tot, g, b))
    print("\ntotal %.3f ms over %d launches (ncu serialises launches and runs them cold-cache: compare shares)" % (tot / 1e6, len(rows)))


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [(w, hdr.index(w)) for w in WANT if w in hdr]
    ki = hdr.index("Kernel Name")
    print("| kernel | " + " | ".join("%s [%s]" % (w, units[i]) for w, i in cols) + " |")
    print("|---|" + "---:|" * len(cols))
    for r in rows[2:]:
        print("| `%s` | " % r[ki].split("(")[0].replace("void ", "") + " | ".join(r[i] for _, i in cols) + " |")


if __name__ == "__main__":
    print("## launch list: `%s`\n" % sys.argv[1])
    launches(sys.argv[1])
    if len(sys.argv) > 2:
        print("\n## `ncu --set full` capture: `%s`\n" % sys.argv[2])
        full(sys.argv[2])
