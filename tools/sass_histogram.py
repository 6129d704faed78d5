"""SASS instruction histogram per kernel of libjpezy_b200.so (cuobjdump -sass): which pipes a kernel leans on and the
mnemonics that prove the Blackwell paths (UBLKCP = cp.async.bulk, SYNCS = mbarrier, FFMA2/FADD2/FMUL2 = packed f32x2,
IDP = dp2a/dp4a, REDUX/CREDUX).  usage: python tools/sass_histogram.py [lib.so] [kernel-name-substring ...]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 and sys.argv[1].endswith(".so") else os.path.join(ROOT, "jpezy_b200", "libjpezy_b200.so")
    want = [a for a in sys.argv[1:] if not a.endswith(".so")]
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    fn, per = None, collections.OrderedDict()
    archs = set()
    for line in out.splitlines():
        m = re.search(r"arch = (sm_\w+)", line)
        if m:
            archs.add(m.group(1))
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = m.group(1)
            per[fn] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and fn:
            per[fn][m.group(1)] += 1
    print("cubins:", ", ".join(sorted(archs)))
    keys = ["UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "FFMA2", "FADD2", "FMUL2", "IDP", "REDUX", "CREDUX", "FMNMX3", "DADD", "DMUL", "DFMA",
            "F2I", "I2F", "I2FP", "LDS", "STS", "LDG", "STG", "ATOMS", "SHFL", "BAR", "HMMA", "UTCHMMA"]
    for f, c in per.items():
        name = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip().split("(")[0]
        if want and not any(w in name for w in want):
            continue
        tot = sum(c.values())
        print("== %s: %d instructions" % (name, tot))
        print("   " + "  ".join("%s %d" % (k, c[k]) for k in keys if c[k]))
        print("   top: " + ", ".join("%s %d" % kv for kv in c.most_common(12)))


if __name__ == "__main__":
    main()
