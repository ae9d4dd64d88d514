"""Histogram of executed warp instructions per SASS opcode from an `ncu --page source --print-source cuda,sass --csv` dump."""
import csv
import collections
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    hdr = rows[h]
    ii = hdr.index("Instructions Executed")
    hist = collections.Counter()
    seen = set()
    for r in rows[h + 1:]:
        if len(r) <= ii or not r[2].startswith("0x") or r[2] in seen:
            continue
        seen.add(r[2])
        try:
            n = int(r[ii] or 0)
        except ValueError:
            continue
        toks = r[3].strip().split()
        if toks and toks[0].startswith("@"):
            toks = toks[1:]
        if toks:
            hist[toks[0].rstrip(";")] += n
    tot = sum(hist.values()) or 1
    print("total", tot)
    for op, n in hist.most_common(top):
        print("%6.2f%%  %12d  %s" % (100.0 * n / tot, n, op))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
