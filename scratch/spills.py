"""Count LDL/STL per source-line range in a cubin: python scratch/spills.py file.o 'kernel substr' lo-hi [lo-hi ...]"""
import re, subprocess, sys, collections, tempfile, os, glob
obj, ranges = sys.argv[1], [tuple(map(int, r.split("-"))) for r in sys.argv[2:]]
d = tempfile.mkdtemp(); subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
sass = subprocess.run(["nvdisasm", "-g", "-c", glob.glob(d + "/*.cubin")[0]], capture_output=True, text=True).stdout
cur = ("?", 0); cnt = collections.Counter(); tot = 0; infn = False
for l in sass.splitlines():
    if re.match(r'^\$_Z\S*\$_Z\S*:', l): break  # first subroutine (cold function) after the kernel body
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m and re.search(r'\b(LDL|STL)(\.|\b)', m.group(2)): cnt[cur] += 1; tot += 1
print("total local ld/st:", tot)
for lo, hi in ranges:
    print(f"  chain_engine.cuh {lo}-{hi}:", sum(v for (f, ln), v in cnt.items() if f == "chain_engine.cuh" and lo <= ln <= hi))
