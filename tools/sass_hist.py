"""Opcode histogram of one kernel from an .ncu-rep source page (samples, instructions, active threads per instruction).
    python tools/sass_hist.py prof.ncu-rep <kernel-regex> [launch-index]"""
import collections
import csv
import re
import subprocess
import sys


def main(path, kernel, which=0):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for r in csv.reader(out.splitlines()):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": [], "hdr": None}
            blocks.append(cur)
            continue
        if cur is None:
            continue
        if cur["hdr"] is None:
            cur["hdr"] = r
            continue
        cur["rows"].append(r)
    b = blocks[which]
    h = b["hdr"]
    iS, iN, iI, iT = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
    ops = collections.defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for r in b["rows"]:
        if len(r) < len(h):
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS])
        op = m.group(2) if m else "?"
        op = ".".join(op.split(".")[:2]) if op.startswith(("MUFU", "LDG", "LDL", "STL", "STG")) else op.split(".")[0]
        for k, i in enumerate((iN, iI, iT)):
            v = int(r[i] or 0)
            ops[op][k] += v
            tot[k] += v
    print(f"# {b['name'][:100]}")
    print(f"# samples {tot[0]}, warp instructions {tot[1]}, thread instructions {tot[2]}, avg active threads {tot[2] / max(tot[1], 1):.2f}, SASS lines {len(b['rows'])}")
    for op, v in sorted(ops.items(), key=lambda x: -x[1][1])[:28]:
        print(f"{op:14s} samples {v[0]:8d} {v[0] / max(tot[0], 1) * 100:5.1f}%   inst {v[1]:11d} {v[1] / max(tot[1], 1) * 100:5.1f}%   threads/inst {v[2] / max(v[1], 1):5.1f}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
