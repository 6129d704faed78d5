"""Summarise `ncu --page source --csv --print-source cuda,sass` output: executed warp instructions and stall
samples per source region (regions = functions/markers found in the source file) and the hottest lines.
usage: python tools/ncu_src_summary.py report.ncu-rep source_file npixels [marker ...]"""
import collections
import csv
import subprocess
import sys


def main():
    rep, srcfile, npx = sys.argv[1], sys.argv[2], float(sys.argv[3])
    markers = sys.argv[4:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    src = open(srcfile).read().split("\n")
    marks = []
    for m in markers:
        hits = [i + 1 for i, l in enumerate(src) if m in l]
        if hits:
            marks.append((m[:40], hits[-1] if m.startswith("//") else hits[0]))
    marks.sort(key=lambda x: x[1])
    per = collections.defaultdict(lambda: [0, 0])
    hdr = None
    for r in rows:
        if r and r[0] == "Line No" and len(r) > 8:
            hdr = r
            iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or len(r) <= iI or not r[0].isdigit() or not r[iI].isdigit():
            continue
        per[int(r[0])][0] += int(r[iI])
        per[int(r[0])][1] += int(r[iS]) if r[iS].isdigit() else 0
    tot = sum(v[0] for v in per.values())
    tots = sum(v[1] for v in per.values()) or 1
    print("total warp-instructions %d  = %.1f thread-slots per pixel" % (tot, tot * 32 / npx))
    agg = collections.OrderedDict((m[0], [0, 0]) for m in marks)
    for ln, (n, s) in per.items():
        name = None
        for nm, st in marks:
            if ln >= st:
                name = nm
        if name:
            agg[name][0] += n
            agg[name][1] += s
    for k, (n, s) in agg.items():
        print("%-42s %9d %5.1f%%  slots/px %6.2f  stall-samples %5.1f%%" % (k, n, 100 * n / tot, n * 32 / npx, 100 * s / tots))
    print("hottest lines:")
    for ln, (n, s) in sorted(per.items(), key=lambda kv: -kv[1][0])[:30]:
        print("%5d %9d %4.1f%% samp %4.1f%% | %s" % (ln, n, 100 * n / tot, 100 * s / tots, src[ln - 1].strip()[:105] if ln <= len(src) else "?"))


if __name__ == "__main__":
    main()
