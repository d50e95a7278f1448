"""Per-source-line instruction counts / stall samples: joins an ncu report's SASS page with nvdisasm --print-line-info.
usage: python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <object.o> <mangled kernel substring> [topN]"""
import collections, csv, io, re, subprocess, sys, tempfile, os
rep, kre, obj, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
blk = '"Kernel Name"' + out.split('"Kernel Name"')[1]
rows = list(csv.reader(io.StringIO(blk)))
hdr = rows[1]
ie, smp, src = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
prof = []
for r in rows[2:]:
    try: prof.append((int(r[ie]), int(r[smp]), r[src]))
    except (ValueError, IndexError): pass
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
lines, cur, inside = [], None, False
for l in sass.splitlines():
    if l.startswith(".text."):
        inside = mangled in l
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): lines.append(cur)
print("profile instrs", len(prof), "disasm instrs", len(lines))
agg = collections.defaultdict(lambda: [0, 0])
for (n, s, _), loc in zip(prof, lines):
    agg[loc][0] += n; agg[loc][1] += s
tot_n = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values())
srcs = {}
def text(loc):
    f, ln = loc
    for base in ("vector_line_quantization_b200/csrc",):
        p = os.path.join(base, f)
        if os.path.exists(p):
            if p not in srcs: srcs[p] = open(p).read().splitlines()
            return srcs[p][ln - 1].strip()[:90]
    return ""
print("%-22s %8s %6s %7s %6s  source" % ("file:line", "winstr", "%", "samples", "%"))
for loc, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if loc is None: continue
    print("%-22s %8d %5.1f%% %7d %5.1f%%  %s" % ("%s:%d" % loc, n, 100 * n / tot_n, s, 100 * s / max(tot_s, 1), text(loc)))
