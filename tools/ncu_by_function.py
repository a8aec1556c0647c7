"""Executed warp instructions, stall samples and static / touched code size of a .ncu-rep per C++ function (line table of the
matching object file): which functions make up the instruction working set.
   python tools/ncu_by_function.py report.ncu-rep nuts_rs_b200/csrc/build/engine_32_4_16_2.o [top]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, obj = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]; ix = {c: i for i, c in enumerate(h)}
body = [r for r in rows[2:] if len(r) >= len(h)]
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, capture_output=True)
    cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cubin)], capture_output=True, text=True).stdout
src_dir = os.path.join(os.path.dirname(os.path.abspath(obj)), "..")
funcs = {}
def load(f):
    p = os.path.join(src_dir, f)
    if f in funcs: return
    o = []
    if os.path.exists(p):
        for n, l in enumerate(open(p), 1):
            m = re.match(r"\s*(?:template\s*<[^>]*>\s*)?(?:static\s+)?__device__\s+(?:__forceinline__\s+|__noinline__\s+|inline\s+)?[\w:<>,\s\*&]*?\b(\w+)\s*\(", l)
            if m and m.group(1) not in ("if", "for", "while", "defined"): o.append((n, m.group(1)))
            elif "__global__" in l:
                m = re.search(r"\b(nuts_\w+)\s*\(", l)
                if m: o.append((n, m.group(1)))
    funcs[f] = o
lines = []; cur = None
for l in dis.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): lines.append(cur)
n = min(len(lines), len(body))
st = collections.defaultdict(lambda: [0, 0, 0, 0])  # static, touched, executed, samples
for (cur, r) in zip(lines[:n], body[:n]):
    if cur is None: key = "?"
    else:
        f, ln = cur; load(f); name = f
        for s, nm in funcs.get(f, []):
            if s <= ln: name = nm
            else: break
        key = name
    ex = int(r[ix["Instructions Executed"]] or 0); sm = int(r[ix["# Samples"]] or 0)
    a = st[key]; a[0] += 1; a[1] += 1 if ex else 0; a[2] += ex; a[3] += sm
tex = sum(a[2] for a in st.values()); tsm = sum(a[3] for a in st.values())
print(f"{'function':34s} {'static':>7s} {'touched':>7s} {'exec %':>7s} {'stall %':>7s}")
for k, a in sorted(st.items(), key=lambda kv: -kv[1][3])[:top]:
    print(f"{k:34s} {a[0]:7d} {a[1]:7d} {100.0*a[2]/max(tex,1):7.2f} {100.0*a[3]/max(tsm,1):7.2f}")
print(f"{'total':34s} {sum(a[0] for a in st.values()):7d} {sum(a[1] for a in st.values()):7d}")
