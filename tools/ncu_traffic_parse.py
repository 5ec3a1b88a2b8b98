#!/usr/bin/env python
"""ncu launch list (csv, metrics gpu__time_duration.sum + dram__bytes_{read,write}.sum) + the phase list tools/ncu_traffic_driver.py printed
-> profiles/traffic.json (what bench.py's roofline.traffic reads).   usage: ncu_traffic_parse.py launches.csv phases.json out.json [commit]"""
import csv, json, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if r]
i = next(k for k, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[i]; data = rows[i + 1:]
iid, iname, imet, iunit, ival = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}
launches = {}
for r in data:
    d = launches.setdefault(int(r[iid]), {"name": r[iname].split("(")[0].replace("void ", "").replace("ssm::", "")})
    d[r[imet]] = float(r[ival].replace(",", "")) * scale.get(r[iunit], 1.0)
ours = [launches[k] for k in sorted(launches)]      # the library's launches, in order (the driver launches nothing else but memsets / copies)
phases = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
out = {"source": f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, tools/ncu_traffic_driver.py; {sys.argv[1]}",
       "commit": sys.argv[4] if len(sys.argv) > 4 else None, "phases": {}}
for ph in phases:
    ks = ours[ph["first_launch"]: ph["first_launch"] + ph["launches"]]
    out["phases"][ph["phase"]] = [{"kernel": k["name"][:90], "ms": k.get("gpu__time_duration.sum"), "dram_bytes": k.get("dram__bytes_read.sum", 0) + k.get("dram__bytes_write.sum", 0)} for k in ks]
tot = lambda name: sum(k["dram_bytes"] for k in out["phases"][name])
out["ab_qr_kernel_dram_bytes_per_launch"] = tot("c2 qr")
out["ab_ldiv_dram_bytes_per_ldiv"] = tot("c2 ldiv")
out["ab_fused_dram_bytes_per_solve"] = tot("c2 fused")
out["bps_qr_dram_bytes_per_qr"] = tot("c3 qr")
out["bps_solve_dram_bytes_per_solve"] = tot("c3 solve")
out["abc_solve_dram_bytes_per_solve"] = tot("c4 solve")
json.dump(out, open(sys.argv[3], "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k.endswith(("launch", "ldiv", "solve", "qr"))}, indent=1))
