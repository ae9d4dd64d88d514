"""Summarise an `ncu --page source --print-source cuda,sass --csv` dump: instructions and stall samples per CUDA line."""
import csv
import sys


def main(path, top=30):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    hdr = rows[h]
    ii, sa = hdr.index("Instructions Executed"), hdr.index("# Samples")
    data = []
    for r in rows[h + 1:]:
        if len(r) <= max(ii, sa) or r[2] != "-":          # keep the per-CUDA-line aggregate rows (Address == "-")
            continue
        try:
            data.append((int(r[ii] or 0), int(r[sa] or 0), r[0], r[1]))
        except ValueError:
            pass
    tot = sum(d[0] for d in data) or 1
    tots = sum(d[1] for d in data) or 1
    print("total warp instructions %d, stall samples %d" % (tot, tots))
    for d in sorted(data, key=lambda x: -x[1])[:top]:
        print("%5.1f%% instr %5.1f%% samples  L%-4s %s" % (d[0] / tot * 100, d[1] / tots * 100, d[2], d[3].strip()[:120]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
