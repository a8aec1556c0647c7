"""SASS evidence for the memory path of the built engines: counts of the wide / asynchronous / cluster instructions per kernel.
   python tools/sass_listing.py > profiles/r2_sass_memory_instructions.txt"""
import collections, glob, os, re, subprocess
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nuts_rs_b200", "csrc", "build")
objs = ["engine_64_16_54_1.o", "engine_32_4_21_2.o", "engine_32_1_16_3.o", "engine_v2_480_21_171.o", "engine_cl_1024_10_41_2.o"]
pats = collections.OrderedDict([
    ("LDG.E.128 (16-byte global load)", r"\bLDG\.E\.128"), ("LDG.E.64", r"\bLDG\.E\.64"), ("STG.E.128 (16-byte global store)", r"\bSTG\.E\.128"),
    ("STG.E.64", r"\bSTG\.E\.64"), ("LDS.128 (16-byte shared load)", r"\bLDS\.128"), ("STS.128", r"\bSTS\.128"),
    ("LDGSTS (cp.async global->shared)", r"\bLDGSTS"), ("LDGDEPBAR / DEPBAR (cp.async groups)", r"\bLDGDEPBAR|\bDEPBAR"),
    ("CCTL.E.PF2 (prefetch.global.L2)", r"\bCCTL\.E\.PF"), ("UCGABAR_ARV / UCGABAR_WAIT (barrier.cluster)", r"UCGABAR"),
    ("ST.E.64 [Rn] without descriptor (st.shared::cluster.f64 to a MAPA address: DSMEM exchange)", r"\bST\.E\.64 \[R\d+\]"),
    ("STG.E.EF.* (evict-first streaming stores: draws, estimator planes)", r"\bSTG\.E\.EF"), ("LDG.E.EF.* (evict-first loads)", r"\bLDG\.E\.EF|\bLD\.E\.EF"),
    ("LD.E.128 / ST.E.128 (16-byte accesses of the cold functions)", r"\bLD\.E\.128|\bST\.E\.128"), ("BAR.SYNC", r"\bBAR\.SYNC"),
    ("SHFL", r"\bSHFL"), ("DFMA", r"\bDFMA"), ("MUFU.RCP64H / RSQ64H", r"MUFU\.(RCP64H|RSQ64H)"), ("LDL (local load)", r"\bLDL"), ("STL (local store)", r"\bSTL"),
])
for o in objs:
    p = os.path.join(root, o)
    if not os.path.exists(p):
        print(f"{o}: not built\n"); continue
    sass = subprocess.run(["cuobjdump", "-sass", p], capture_output=True, text=True).stdout
    kern = re.findall(r"Function : (\S+)", sass)
    demangled = subprocess.run(["cu++filt"] + kern[:1], capture_output=True, text=True).stdout.strip()
    lines = [l for l in sass.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l)]
    print(f"{o}: {demangled}\n  {len(lines)} SASS instructions ({len(lines) * 16 // 1024} KB)")
    for name, pat in pats.items():
        n = sum(1 for l in lines if re.search(pat, l))
        print(f"  {n:6d}  {name}")
    ex = [l.strip()[:110] for l in lines if re.search(r"LDG\.E\.128|STG\.E\.128|LDGSTS|CCTL\.E\.PF|UCGABAR", l)]
    seen = set(); shown = 0
    for l in ex:
        k = re.sub(r"R\d+|UR\d+|0x[0-9a-f]+|/\*[0-9a-f]+\*/", "", l)
        if k in seen: continue
        seen.add(k); shown += 1
        print("      e.g. " + l)
        if shown >= 6: break
    print()
