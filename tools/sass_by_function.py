"""SASS instruction count of an engine object per C++ function (by the line table): where the code size of a kernel comes from.
   python tools/sass_by_function.py nuts_rs_b200/csrc/build/engine_32_4_16_2.o [top]"""
import collections, os, re, subprocess, sys, tempfile
obj = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, capture_output=True)
    cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cubin)], capture_output=True, text=True).stdout
src_dir = os.path.join(os.path.dirname(os.path.abspath(obj)), "..")
funcs = {}  # file -> sorted [(start line, name)]
def load(f):
    p = os.path.join(src_dir, f)
    if f in funcs or not os.path.exists(p): return
    out = []
    for n, l in enumerate(open(p), 1):
        m = re.match(r"\s*(?:template\s*<[^>]*>\s*)?(?:static\s+)?__device__\s+(?:__forceinline__|__noinline__|inline)?\s*[\w:<>,\s\*&]*?\b(\w+)\s*\(", l)
        if m and m.group(1) not in ("if", "for", "while"): out.append((n, m.group(1)))
        m = re.match(r"\s*__global__\s+void.*?\b(\w+)\s*\(", l) or re.match(r"__global__ void \w+\(.*?\)\s*(\w+)\(", l)
        if m: out.append((n, m.group(1)))
    funcs[f] = out
cur = None; cnt = collections.Counter(); total = 0
for l in dis.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        total += 1
        if cur is None: cnt[("?", "?")] += 1; continue
        f, ln = cur; load(f)
        name = "?"
        for s, n in funcs.get(f, []):
            if s <= ln: name = n
            else: break
        cnt[(f, name)] += 1
print(f"{total} SASS instructions = {total * 16 / 1024:.0f} KB")
for (f, n), c in cnt.most_common(top):
    print(f"{c:7d} {100.0 * c / total:5.1f}%  {f}:{n}")
