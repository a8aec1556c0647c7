"""Turn a .ncu-rep of the draw kernel into the text + json summaries committed under profiles/."""
import collections, csv, io, json, subprocess, sys
rep, out_prefix = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static"]
lines = [f"ncu --set full --clock-control none, report {rep}", ""]
summary = {}
for k in keys:
    if k in m:
        lines.append(f"{k:70s} {m[k][0]} {m[k][1]}")
        summary[k] = m[k][0]
def gb(x):
    v, u = m[x]
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
traffic = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
summary["dram_traffic_bytes_per_launch"] = traffic
lines += ["", f"DRAM traffic per launch (read + write): {traffic/1e9:.3f} GB", ""]
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
h = rows[1]; data = [r for r in rows[2:] if len(r) >= len(h)]
ix = {x: i for i, x in enumerate(h)}
ops = collections.Counter(); stalls = collections.Counter(); tot = 0
for r in data:
    n = int(r[ix["Instructions Executed"]] or 0); tot += n
    op = r[ix["Source"]].strip().split(); o = op[1] if op[0].startswith("@") else op[0]
    ops[o.split(".")[0]] += n
    for c in h:
        if c.startswith("stall_") and "Not Issued" not in c and r[ix[c]]:
            stalls[c] += int(r[ix[c]])
lines.append(f"SASS instructions in kernel: {len(data)}; warp instructions executed: {tot}")
lines.append("instruction mix: " + ", ".join(f"{o} {100*n/tot:.1f}%" for o, n in ops.most_common(14)))
st = sum(stalls.values())
lines.append("warp stall samples: " + ", ".join(f"{c[6:]} {100*n/st:.1f}%" for c, n in stalls.most_common(8)))
summary["instruction_mix_pct"] = {o: round(100 * n / tot, 2) for o, n in ops.most_common(14)}
summary["stall_pct"] = {c[6:]: round(100 * n / st, 2) for c, n in stalls.most_common(8)}
open(out_prefix + ".txt", "w").write("\n".join(lines) + "\n")
json.dump(summary, open(out_prefix + ".json", "w"), indent=1)
print("\n".join(lines))
