"""Aggregate an ncu source-page profile by C++ function: python scratch/agg_profile.py rep.ncu-rep file.o
Function ranges are taken from the sources (a function = from its definition line to the next definition)."""
import csv, re, subprocess, sys, collections, tempfile, os, glob, io
rep, obj = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__))) + "/nuts_rs_b200/csrc/"
ranges = {}
for f in ["chain_engine.cuh", "chain_engine_v2.cuh", "device_common.cuh"]:
    starts = []
    for n, l in enumerate(open(root + f).read().splitlines(), 1):
        m = re.match(r'\s*(?:static\s+)?(?:template\s*<[^>]*>\s*)?__(?:device|global|host)__.*?\b(\w+)\s*\([^;]*$', l)
        if m and not l.strip().startswith("//"): starts.append((n, m.group(1)))
    ranges[f] = starts
def fn_of(f, ln):
    if f not in ranges: return f
    name = f
    for n, nm in ranges[f]:
        if n <= ln: name = nm
        else: break
    return name
d = tempfile.mkdtemp(); subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
dis = subprocess.run(["nvdisasm", "-g", "-c", glob.glob(d + "/*.cubin")[0]], capture_output=True, text=True).stdout
inst = []; cur = ("?", 0)
for l in dis.splitlines():
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: inst.append((m.group(2).strip(), cur))
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass))); hdr = rows[1]; data = [r for r in rows[2:] if len(r) >= len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
assert len(inst) == len(data), (len(inst), len(data))
stall_cols = [c for c in hdr if c.startswith("stall_")]
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for (op, (f, ln)), r in zip(inst, data):
    k = fn_of(f, ln); a = agg[k]
    a[0] += int(r[ix["# Samples"]] or 0); a[1] += int(r[ix["Instructions Executed"]] or 0)
    for c in stall_cols:
        if r[ix[c]]: a[2][c[6:]] += int(r[ix[c]])
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
print(f"total samples {ts}, warp instructions {ti}, static instructions {len(inst)}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    top = ", ".join(f"{n} {100*c/max(1,sum(v[2].values())):.0f}%" for n, c in v[2].most_common(3))
    print(f"{k:28s} samp {100*v[0]/ts:6.2f}%  inst {100*v[1]/ti:6.2f}%   {top}")
