"""Opcode histogram of an `ncu --page source --print-source cuda,sass --csv` dump with executed counts and the average
number of active (predicated-on) lanes per executed warp instruction; optional: restrict to a CUDA source line range.
    python profiles/ncu_sass_hist.py src.csv [first_line last_line]"""
import collections
import csv
import sys


def main(path, lo=None, hi=None):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    hdr = rows[h]
    ii, ti, pi = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Predicated-On Thread Instructions Executed")
    hist, thr = collections.Counter(), collections.Counter()
    cur = None
    for r in rows[h + 1:]:
        if len(r) <= pi:
            continue
        if r[2] == "-":
            cur = int(r[0]) if r[0].isdigit() else None
            continue
        if lo is not None and (cur is None or cur < lo or cur > hi):
            continue
        try:
            n, t = int(r[ii] or 0), int(r[pi] or 0)
        except ValueError:
            continue
        tok = r[3].split()
        if not tok:
            continue
        op = tok[1] if tok[0].startswith("@") and len(tok) > 1 else tok[0]
        hist[op] += n
        thr[op] += t
    tot = sum(hist.values()) or 1
    print("warp instructions %d, active lanes per instruction %.1f" % (tot, sum(thr.values()) / tot))
    for op, n in hist.most_common(30):
        print("%-22s %11d %5.1f%%  lanes %.1f" % (op, n, n / tot * 100, thr[op] / max(n, 1)))


if __name__ == "__main__":
    a = sys.argv
    main(a[1], int(a[2]) if len(a) > 3 else None, int(a[3]) if len(a) > 3 else None)
