"""Join ncu's per-SASS-instruction profile with nvdisasm line info -> per source line totals."""
import csv, re, subprocess, sys, collections
cubin, sass_csv, kernel_substr = sys.argv[1], sys.argv[2], sys.argv[3]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# keep only the kernel function section: instructions listed in order; device functions follow in the same listing
lines = dis.splitlines()
inst_line = []  # (offset, opcode text, file, line, inlined chain)
cur = ("?", 0)
in_func = False
for l in lines:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        inst_line.append((int(m.group(1), 16), m.group(2).strip(), cur))
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) >= len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
print("nvdisasm instructions", len(inst_line), "ncu instructions", len(data))
# ncu lists kernel function then device functions, nvdisasm too (per .text section); align by opcode sequence greedily
by_line = collections.defaultdict(lambda: [0, 0, 0, 0])  # samples, inst, local inst, long_sb
k = 0
mismatch = 0
for r in data:
    src = r[ix["Source"]].strip()
    op = src.split()[1] if src.startswith("@") else src.split()[0]
    # advance k until opcode matches (handles small differences)
    kk = k
    while kk < len(inst_line) and kk < k + 50:
        t = inst_line[kk][1]
        top = t.split()[1] if t.startswith("@") else t.split()[0]
        if top == op:
            break
        kk += 1
    if kk >= len(inst_line) or kk >= k + 50:
        mismatch += 1
        continue
    k = kk + 1
    f, ln = inst_line[kk][2]
    e = by_line[(f, ln)]
    e[0] += int(r[ix["# Samples"]] or 0)
    n = int(r[ix["Instructions Executed"]] or 0)
    e[1] += n
    if "LDL" in op or "STL" in op:
        e[2] += n
    e[3] += int(r[ix["stall_long_sb"]] or 0)
print("unmatched", mismatch)
tot = sum(e[0] for e in by_line.values()); toti = sum(e[1] for e in by_line.values())
src_cache = {}
def srcline(f, ln):
    try:
        if f not in src_cache:
            src_cache[f] = open("/root/repo/nuts_rs_b200/csrc/" + f).read().splitlines()
        return src_cache[f][ln - 1].strip()[:90]
    except Exception:
        return ""
print(f"{'file:line':28s} {'samp%':>6s} {'inst%':>6s} {'local':>9s} {'longsb%':>7s}  source")
for (f, ln), e in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[4]) if len(sys.argv) > 4 else 45]:
    print(f"{f+':'+str(ln):28s} {100*e[0]/tot:6.2f} {100*e[1]/toti:6.2f} {e[2]:9d} {100*e[3]/max(1,e[0]):7.1f}  {srcline(f, ln)}")
