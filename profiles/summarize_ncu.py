"""Summarise `ncu -i <rep> --page raw --csv` dumps into a markdown table (one row per captured launch).

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python profiles/summarize_ncu.py /tmp/raw.csv [more.csv ...] > profiles/rNN_ncu_summary.md

Columns: duration, DRAM read / written per launch, DRAM throughput % of peak, issue-slot utilisation, achieved
occupancy, registers, grid size, the two largest warp-stall reasons per issued instruction.
`python profiles/summarize_ncu.py --traffic raw.csv ...` prints {kernel: dram bytes per launch} as JSON instead
(bench.py reads profiles/traffic.json for `roofline.traffic`)."""
import csv
import json
import re
import sys

COLS = [("gpu__time_duration.sum", "time"),
        ("dram__bytes_read.sum", "dram rd"),
        ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
        ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"),
        ("smsp__inst_executed.sum", "warp instr")]
STALLS = ["long_scoreboard", "short_scoreboard", "barrier", "mio_throttle", "lg_throttle", "wait", "math_pipe_throttle",
          "not_selected", "no_instruction", "imc_miss", "membar", "sleeping", "drain", "dispatch_stall", "branch_resolving",
          "tex_throttle"]

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("b200seg::", "").replace("void ", "").strip()


def rows_of(path):
    lines = open(path).read().splitlines()
    start = next(i for i, ln in enumerate(lines) if ln.startswith('"ID"'))      # skip the ==PROF== preamble of --log-file
    rows = list(csv.reader(lines[start:]))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        if len(r) == len(hdr):
            yield hdr, units, r


def main(argv):
    traffic = "--traffic" in argv
    paths = [a for a in argv if not a.startswith("--")]
    out_traffic = {}
    if not traffic:
        print("| kernel | " + " | ".join(c[1] for c in COLS) + " | top stalls (warps per issue) |")
        print("|---|" + "---|" * (len(COLS) + 1))
    for path in paths:
        for hdr, units, r in rows_of(path):
            name = short(r[hdr.index("Kernel Name")])
            cells = []
            for key, _ in COLS:
                if key in hdr:
                    i = hdr.index(key)
                    v = r[i]
                    try:
                        f = float(v.replace(",", ""))
                        v = ("%.4g" % f)
                    except ValueError:
                        pass
                    cells.append("%s %s" % (v, units[i]) if units[i] and key not in ("launch__grid_size",) else v)
                else:
                    cells.append("-")
            st = []
            for s in STALLS:
                k = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s
                if k in hdr:
                    try:
                        st.append((float(r[hdr.index(k)]), s))
                    except ValueError:
                        pass
            st.sort(reverse=True)
            if traffic:
                tot = 0.0
                for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    i = hdr.index(key)
                    tot += float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
                out_traffic[name] = tot
            else:
                print("| %s | %s | %s |" % (name, " | ".join(cells), ", ".join("%s %.2f" % (s, v) for v, s in st[:2])))
    if traffic:
        print(json.dumps(out_traffic, indent=1))


if __name__ == "__main__":
    main(sys.argv[1:])
