#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS export with nvdisasm -gi line info, aggregate per source line.

usage: sass_lines.py <ncu_source.csv> <nvdisasm_all.sass> <mangled-substring> [top] [--outer] [--dump] [--regions file.json]

  ncu -i rep.ncu-rep --page source --csv > ncu_source.csv        (a capture with --section SourceCounters)
  cuobjdump -xelf all libb2pt.so; nvdisasm -gi -c *.cubin > nvdisasm_all.sass
  --regions: {"lo": L0, "hi": L1, "ranges": {name: [a, b], ...}} groups the instructions whose inline chain passes through
             pt_math.cuh lines L0..L1 (one walk step) by the line range of that frame; everything else by kernel line.
"""
import csv, re, sys, collections

src_csv, sass, key = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# nvdisasm: find function
lines = open(sass).read().split("\n")
start = None
for i, l in enumerate(lines):
    if l.startswith(".text.") and key in l and l.rstrip().endswith(":"):
        start = i; break
assert start is not None
cur = None
fresh = True
level = 0 if '--outer' not in sys.argv else -1
info = {}  # offset -> (file, line, chain)
for l in lines[start + 1:]:
    if l.startswith("//-----") or (l.startswith(".text.") and l.rstrip().endswith(":")):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        f = m.group(1).split("/")[-1]
        chain = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
        here = (f, int(m.group(2)))
        if fresh: cur = [here]; fresh = False
        else: cur.append(here)
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        info[int(m.group(1), 16)] = (tuple(cur) if cur else None, m.group(2).strip())
        fresh = True
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
col = {n: i for i, n in enumerate(hdr)}
base = None
agg = collections.defaultdict(lambda: [0, 0, 0, 0])  # inst, thread-inst, samples, n_sass
tot = [0, 0, 0]
per_inst = []
for r in rows[2:]:
    if len(r) < len(hdr) - 5: continue
    addr = int(r[col["Address"]], 16)
    if base is None: base = addr
    off = addr - base
    ie = int(r[col["Instructions Executed"]] or 0)
    te = int(r[col["Thread Instructions Executed"]] or 0)
    sm = int(r[col["# Samples"]] or 0)
    where, text = info.get(off, (None, "?"))
    k = where[level] if where else ("?", 0)
    a = agg[k]; a[0] += ie; a[1] += te; a[2] += sm; a[3] += 1
    tot[0] += ie; tot[1] += te; tot[2] += sm
    per_inst.append((off, ie, te, sm, k, text, r))
if "--regions" in sys.argv:
    import json
    cfg = json.load(open(sys.argv[sys.argv.index("--regions") + 1]))
    def region(chain):
        for f, l in chain or ():
            if f == "pt_math.cuh" and cfg["lo"] <= l <= cfg["hi"]:
                for name, (a, b) in cfg["ranges"].items():
                    if a <= l <= b: return name
                return "step:other"
        for f, l in reversed(chain or ()):
            if f == "pt_kernels.cu": return f"kernel:{l}"
        return "other"
    ragg = collections.defaultdict(lambda: [0, 0, 0])
    for off, ie, te, sm, k, text, r in per_inst:
        a = ragg[region(info.get(off, (None, ""))[0])]; a[0] += ie; a[1] += te; a[2] += sm
    print(f"{'region':22s} {'inst%':>6s} {'lanes':>6s} {'smpl%':>6s}")
    for k, a in sorted(ragg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{k:22s} {100*a[0]/tot[0]:6.2f} {a[1]/max(a[0],1):6.2f} {100*a[2]/max(tot[2],1):6.2f}")
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    st = {c[6:]: sum(int(r[col[c]] or 0) for *_, r in per_inst) for c in stall_cols}
    print("stall reasons (% of samples):", {k: round(100 * v / max(tot[2], 1), 1) for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v})
print(f"total warp-inst {tot[0]}  thread-inst {tot[1]}  avg lanes {tot[1]/max(tot[0],1):.2f}  samples {tot[2]}")
print(f"{'file:line':28s} {'inst%':>6s} {'lanes':>6s} {'smpl%':>6s} {'#sass':>5s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]+':'+str(k[1]):28s} {100*a[0]/tot[0]:6.2f} {a[1]/max(a[0],1):6.2f} {100*a[2]/max(tot[2],1):6.2f} {a[3]:5d}")
if "--dump" in sys.argv:
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    for off, ie, te, sm, k, text, r in per_inst:
        if ie == 0: continue
        st = sorted(((int(r[col[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
        print(f"{off:05x} {ie:8d} {te/max(ie,1):5.1f} {sm:6d} {k[0]}:{k[1]:<5d} {text[:60]:60s} {st}")
