"""Stall samples of a .ncu-rep per CUDA source line.  The report's SASS page gives samples per instruction (in kernel order);
the line table comes from nvdisasm -g of the matching object file (same build => same instruction sequence).
   python tools/ncu_hot_lines.py report.ncu-rep nuts_rs_b200/csrc/build/engine_64_16_54_1.o [top]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, obj = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]; ix = {c: i for i, c in enumerate(h)}
sass = [(r[ix["Source"]].strip(), int(r[ix["# Samples"]] or 0), int(r[ix["Instructions Executed"]] or 0)) for r in rows[2:] if len(r) >= len(h)]
stall_cols = [c for c in h if c.startswith("stall_")]
stalls = collections.Counter()
for r in rows[2:]:
    if len(r) >= len(h):
        for c in stall_cols:
            if r[ix[c]]: stalls[c] += int(r[ix[c]])
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, capture_output=True)
    cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cubin)], capture_output=True, text=True).stdout
lines = []  # (file, line) per instruction in order
cur = ("?", 0)
for l in dis.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): lines.append(cur)
n = min(len(lines), len(sass))
if len(lines) != len(sass): print(f"warning: {len(sass)} instructions in the report, {len(lines)} in the object; matching the first {n}")
agg = collections.Counter(); inst = collections.Counter(); tot = 0
why = collections.defaultdict(collections.Counter)
body = [r for r in rows[2:] if len(r) >= len(h)]
for (f, ln), (src, s, ie), r in zip(lines[:n], sass[:n], body[:n]):
    agg[(f, ln)] += s; inst[(f, ln)] += ie; tot += s
    for c in stall_cols:
        if "Not Issued" not in c and r[ix[c]] and r[ix[c]] != "0": why[(f, ln)][c[6:]] += int(r[ix[c]])
print(f"total samples {tot}; stalls: " + ", ".join(f"{k[6:]} {100*v/max(1,sum(stalls.values())):.1f}%" for k, v in stalls.most_common(8)))
srcs = {}
for (f, ln), s in agg.most_common(top):
    if f not in srcs:
        p = os.path.join(os.path.dirname(os.path.abspath(obj)), "..", f)
        alt = os.path.join(os.path.dirname(os.path.abspath(obj)), "..", "..", "..", "include", f)
        srcs[f] = open(p).read().splitlines() if os.path.exists(p) else (open(alt).read().splitlines() if os.path.exists(alt) else [])
    text = srcs[f][ln - 1].strip()[:110] if 0 < ln <= len(srcs[f]) else ""
    top2 = ", ".join(f"{k} {v}" for k, v in why[(f, ln)].most_common(2))
    print(f"{100*s/max(tot,1):6.2f}%  {inst[(f, ln)]:>12}  {f}:{ln}: {text}   [{top2}]")
