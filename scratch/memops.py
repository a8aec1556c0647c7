"""Memory instruction kinds per C++ function in the kernel body of an object file: python scratch/memops.py file.o [n]"""
import re, collections, sys, subprocess, tempfile, os, glob
obj = sys.argv[1]
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
cur = ("?", 0); cnt = collections.defaultdict(collections.Counter); total = 0
for l in dis.splitlines():
    if re.match(r'^\$_Z\S*\$_Z\S*cold', l): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        op = m.group(2).split(); o = (op[1] if op[0].startswith('@') else op[0]).split('.')[0]; total += 1
        if o in ('LD', 'ST', 'LDS', 'STS', 'LDL', 'STL', 'LDG', 'STG'): cnt[fn_of(*cur)][o] += 1
        cnt[fn_of(*cur)]['n'] += 1
print("kernel body instructions:", total)
for k, v in sorted(cnt.items(), key=lambda kv: -kv[1]['n'])[:int(sys.argv[2]) if len(sys.argv) > 2 else 24]: print(f"{k:26s}", dict(v))
