"""DRAM traffic per launch of the search kernels, from an `ncu --set full` capture -> profiles/r01_traffic.json
(read by bench.py for roofline.traffic).  usage: python tools/ncu_traffic.py <report.ncu-rep> <db_vectors_per_gpu> <nq>"""
import collections, csv, json, subprocess, sys

rep, n, nq = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
ki, ri, wi = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
stage = {"l2_tc_kernel": "l2_distances", "coarse_select_lines": "coarse_select_lines", "scan_topk": "scan_topk",
         "scan_async": "scan_topk", "term3": "scan_topk"}
acc, cnt = collections.defaultdict(float), collections.defaultdict(int)
for r in rows[2:]:
    for key, st in stage.items():
        if key in r[ki]:
            acc[st] += float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]
            if key != "term3":
                cnt[st] += 1
json.dump({"source": rep, "db_vectors_per_gpu": n, "nq": nq,
           "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch (scan_topk includes its term-3 table kernel)",
           "bytes_per_launch": {k: acc[k] / max(1, cnt[k]) for k in acc}}, sys.stdout, indent=1)
