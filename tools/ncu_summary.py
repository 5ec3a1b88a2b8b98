#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` into the small JSON summaries kept under profiles/.
usage: ncu_summary.py report.ncu-rep out.json "note" """
import csv, io, json, subprocess, sys
rep, out, note = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
res = {}
def num(d, k):
    try: return float(d[k].replace(",", ""))
    except Exception: return None
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d.get("Kernel Name", "?").replace("void ", "").replace("ssm::", "")
    name = name[: name.index("(")] if "(" in name else name
    k = name; i = 2
    while k in res: k = f"{name} #{i}"; i += 1
    dur = num(d, "gpu__time_duration.sum"); rd = num(d, "dram__bytes_read.sum"); wr = num(d, "dram__bytes_write.sum")
    unit = dict(zip(hdr, rows[1]))
    def gb(v, key):
        u = unit.get(key, "")
        return None if v is None else v * {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9, "Tbyte": 1e3}.get(u, 1.0)
    durms = None if dur is None else dur * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(unit.get("gpu__time_duration.sum", "ms"), 1.0)
    e = {"duration_ms": durms, "dram_read_GB": gb(rd, "dram__bytes_read.sum"), "dram_write_GB": gb(wr, "dram__bytes_write.sum"),
         "grid": num(d, "launch__grid_size"), "block": num(d, "launch__block_size"), "regs": num(d, "launch__registers_per_thread"),
         "dyn_smem_per_block": num(d, "launch__shared_mem_per_block_dynamic"),
         "fp64_pipe_pct": num(d, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
         "warp_insts": num(d, "smsp__inst_executed.sum"), "issue_active_pct": num(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
         "warps_active_per_scheduler": num(d, "smsp__warps_active.avg.per_cycle_active"),
         "warp_latency_per_inst": num(d, "smsp__average_warp_latency_per_inst_issued.ratio")}
    for s in ("wait", "short_scoreboard", "long_scoreboard", "barrier", "math_pipe_throttle", "branch_resolving", "not_selected"):
        e["stall_" + s] = num(d, f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio")
    if e["duration_ms"] and e["dram_read_GB"] is not None:
        e["dram_GBps"] = (e["dram_read_GB"] + e["dram_write_GB"]) / (e["duration_ms"] * 1e-3)
    res[k] = e
res["_note"] = note
json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k: {kk: v[kk] for kk in ("duration_ms", "dram_GBps", "regs", "fp64_pipe_pct", "issue_active_pct", "warps_active_per_scheduler")} for k, v in res.items() if k != "_note"}, indent=1))
