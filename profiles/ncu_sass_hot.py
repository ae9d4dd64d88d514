"""Top SASS instructions of an `ncu --page source --print-source cuda,sass --csv` dump by stall samples (with the
dominant stall reason), in program order.   python profiles/ncu_sass_hot.py src.csv [n]"""
import csv
import sys


def num(s):
    try:
        return int(s)
    except ValueError:
        return 0


def main(path, n=30):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    hdr = rows[h]
    si, ii, wi = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    stall_cols = [(i, c) for i, c in enumerate(hdr) if c.startswith("stall_")]
    sass = [r for r in rows[h + 1:] if len(r) > si and r[2] != "-"]
    vals = [num(r[si]) for r in sass]
    tot = sum(vals) or 1
    print("total samples", tot, "| stall columns:", len(stall_cols))
    best = sorted(range(len(vals)), key=lambda i: -vals[i])[:n]
    for i in sorted(best):
        r = sass[i]
        why = ""
        if stall_cols:
            top = sorted(((num(r[c]), nm) for c, nm in stall_cols if c < len(r)), reverse=True)[:2]
            why = " ".join("%s=%d" % (nm.replace("stall_", ""), v) for v, nm in top if v)
        print("%5d smp=%5d (%4.1f%%) ex=%9s  %-70s %s" % (i, vals[i], vals[i] / tot * 100, r[ii], r[3].strip()[:70], why))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
