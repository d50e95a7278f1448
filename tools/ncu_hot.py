"""usage: python tools/ncu_hot.py <report.ncu-rep> <kernel regex> [topN] -- hottest SASS lines with stall reasons + metrics"""
import csv, subprocess, sys, io
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name"')
blk = '"Kernel Name"' + blocks[1]
rows = list(csv.reader(io.StringIO(blk)))
print(rows[0][1][:100])
hdr = rows[1]
src, smp = hdr.index("Source"), hdr.index("# Samples")
names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
idx = {n: hdr.index(n) for n in names}
data = []
tot = 0
for r in rows[2:]:
    if len(r) <= smp: continue
    try: n = int(r[smp])
    except ValueError: continue
    tot += n; data.append((n, r))
agg = {n: 0 for n in names}
for n, r in data:
    for k, i in idx.items():
        try: agg[k] += int(r[i])
        except ValueError: pass
print("samples", tot, {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.01})
data.sort(key=lambda t: -t[0])
for n, r in data[:top]:
    st = {k.replace("stall_", ""): r[i] for k, i in idx.items() if r[i] not in ("0", "")}
    print("%6d %-70s %s" % (n, r[src][:70], st))
