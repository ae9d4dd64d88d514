"""Static SASS opcode histograms of the hottest kernels (cuobjdump -sass on the built objects; no GPU needed).

    python profiles/sass_opcode_hist.py > profiles/r2_sass_opcodes.md

What to look for (profiling guide, "What proves a Blackwell-native kernel"): this hot path is HBM / issue bound integer and
byte work, so there is no tensor-core mnemonic to expect; the sm_100a-specific ones that do show are FFMA2 (fma.rn.f32x2),
VIMNMX3 / VIMNMX3.U16x2 (three-input integer min / max), ATOMS.POPC.INC (aggregated shared-memory increments), REDUX,
128-bit LDG / STG with .NA / .CONSTANT qualifiers."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
OBJ = os.path.join(ROOT, "instanceseg-without-voxelwise-labeling_b200", "build")
KERNELS = [("paste.o", "paste_labels_kernel"), ("soma_binarize.o", "soma_binarize_kernel"), ("peaks3d.o", "peaks_scan3w_kernelILb1ELi3"),
           ("peaks3d.o", "peaks_collect_kernel"), ("peaks3d.o", "peaks_finalize_kernel"), ("largest_cc.o", "largest_cc_fill_kernel"),
           ("roialign3d.o", "roialign3d_fwd_kernelIfLi8"), ("roialign3d.o", "roialign3d_bwd_kernelIfLi8ELi7"), ("nms3d.o", "nms_mask_kernel"),
           ("iou3d.o", "iou3d_kernel")]
SPECIAL = ["FFMA2", "VIMNMX3", "VIMNMX", "ATOMS.POPC.INC", "ATOMS", "REDUX", "LDG.E.128", "STG.E.128", "LDGSTS", "UBLKCP", "UTMALDG", "HMMA", "UTC",
           "DFMA", "DMUL", "DADD", "SHFL", "VOTE", "MATCH", "PRMT", "LDS.128", "BAR"]


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, body = None, collections.defaultdict(list)
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and cur:
            body[cur].append(m.group(1).strip())
    return body


def main():
    cache = {}
    print("# Round 2: static SASS opcode mix of the hottest kernels (`profiles/sass_opcode_hist.py`, cuobjdump -sass, sm_100a)\n")
    for obj, pat in KERNELS:
        path = os.path.join(OBJ, obj)
        if path not in cache:
            cache[path] = functions(path)
        names = [n for n in cache[path] if pat in n]
        if not names:
            print("## %s: not found in %s\n" % (pat, obj))
            continue
        name = sorted(names, key=len)[0]
        ins = cache[path][name]
        ops = collections.Counter()
        full = collections.Counter()
        for i in ins:
            t = i.split()
            op = t[1] if t[0].startswith("@") else t[0]
            ops[op.split(".")[0]] += 1
            full[op] += 1
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()[:110]
        print("## `%s`  (%d instructions)\n" % (demangled, len(ins)))
        print("top opcodes: " + ", ".join("%s %d" % kv for kv in ops.most_common(14)) + "\n")
        sp = []
        for k in SPECIAL:
            c = sum(v for f, v in full.items() if f.startswith(k))
            if c:
                sp.append("%s %d" % (k, c))
        print("mnemonics of interest: " + (", ".join(sp) if sp else "-") + "\n")


if __name__ == "__main__":
    main()
