#!/usr/bin/env python
"""SASS census of the built CUDA library: per kernel, how many of the Blackwell-specific instructions it contains
(B200_PROFILING.md, "What proves a Blackwell-native kernel").  Runs without a GPU: `cuobjdump -sass` on every object of
vector_line_quantization_b200/build/obj (the objects lib/libvlq_b200.so is linked from).

  python tools/sass_census.py > profiles/r02_sass_census.txt
"""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = [("tcgen05.mma", r"\bUTC[A-Z]*MMA\b"), ("tcgen05.ld/st", r"\b(LDTM|STTM)\b"), ("tcgen05 barrier/alloc", r"\bUTC(BAR|ATOMSWS|CP)\b"),
       ("cp.async.bulk (TMA copy)", r"\bUBLKCP\b"), ("cp.async.bulk.prefetch.L2", r"\bUBLKPF\b"), ("TMA tensor", r"\bUTMA(LDG|STG)\b"),
       ("mbarrier", r"\bSYNCS\b"), ("cp.async (LDGSTS)", r"\bLDGSTS\b"), ("legacy HMMA", r"\bHMMA\b"),
       ("LDG.128", r"\bLDG\.E(\.[A-Z0-9]+)*\.128\b"), ("LDS", r"\bLDS\b"), ("ATOMS", r"\bATOMS\b")]


def demangle(names):
    if not names:
        return names
    try:
        out = subprocess.run(["c++filt"] + names, capture_output=True, text=True, stdin=subprocess.DEVNULL,
                             timeout=60).stdout.split("\n")
        return [o if o else n for o, n in zip(out, names)]
    except Exception:
        return names


def main():
    objs = sorted(glob.glob(os.path.join(ROOT, "vector_line_quantization_b200", "build", "obj", "*.o")))
    if not objs:
        sys.exit("build first: python -m vector_line_quantization_b200.build")
    print("# SASS census (cuobjdump -sass, sm_100a objects of lib/libvlq_b200.so); counts are static instruction counts per kernel")
    print("# columns: " + " | ".join(p[0] for p in PAT))
    for o in objs:
        sass = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True, stdin=subprocess.DEVNULL,
                              timeout=300).stdout
        kernels = re.split(r"\n\s*Function : ", sass)[1:]
        print("\n## %s  (%d kernels)" % (os.path.basename(o), len(kernels)))
        names = [k.split("\n", 1)[0].strip() for k in kernels]
        for name, body in zip(demangle(names), kernels):
            counts = [len(re.findall(p, body)) for _, p in PAT]
            ninstr = len(re.findall(r"^\s+/\*[0-9a-f]{4,}\*/", body, flags=re.M))
            short = re.sub(r"\(.*", "", name)
            print("%-78s instrs %6d | %s" % (short[:78], ninstr, " | ".join("%4d" % c for c in counts)))


if __name__ == "__main__":
    main()
