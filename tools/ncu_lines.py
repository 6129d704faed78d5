"""Executed warp instructions and stall samples per source line of an ncu report captured with --import-source on.
usage: python tools/ncu_lines.py report.ncu-rep [npixels] [top] [kernel regex]     (source text is read from the working tree)"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep = sys.argv[1]
    npx = float(sys.argv[2]) if len(sys.argv) > 2 else 3840 * 2160
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
    if len(sys.argv) > 4:
        cmd += ["-k", "regex:" + sys.argv[4]]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    cur, hdr = None, None
    per = collections.defaultdict(lambda: [0, 0, 0])
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur = os.path.basename(r[1])
        elif r[0] == "Line No":
            hdr = r
            iI, iS, iT = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
        elif hdr and r[0].isdigit() and len(r) > iT and r[iI].isdigit():
            v = per[(cur, int(r[0]))]
            v[0] += int(r[iI])
            v[1] += int(r[iS]) if r[iS].isdigit() else 0
            v[2] += int(r[iT]) if r[iT].isdigit() else 0
    tot = sum(v[0] for v in per.values())
    tots = sum(v[1] for v in per.values()) or 1
    print("warp instructions %d = %.1f issue slots (x32) per pixel; thread instructions per pixel %.1f" % (tot, tot * 32 / npx, sum(v[2] for v in per.values()) / npx))
    src = {}
    for (f, ln), (n, s, tn) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        if f not in src:
            pth = os.path.join(ROOT, "jpezy_b200", "csrc", f)
            src[f] = open(pth).read().split("\n") if os.path.exists(pth) else []
        line = src[f][ln - 1].strip()[:96] if ln <= len(src[f]) else "?"
        print("%-20s %4d %9d %5.1f%% samp %5.1f%% thr %4.1f | %s" % (f[:20], ln, n, 100 * n / tot, 100 * s / tots, tn / max(n, 1), line))


if __name__ == "__main__":
    main()
