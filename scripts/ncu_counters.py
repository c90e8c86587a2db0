"""profiles/kernel_counters.json from ncu metric passes (CSV written by `ncu --csv --log-file`, one row per launch and metric):
   python scripts/ncu_counters.py <key> <launches.csv> <kernel-name regex> [steps captured] [note]
sums smsp__inst_executed.sum, dram__bytes_read.sum + dram__bytes_write.sum and gpu__time_duration.sum over the matching
launches (divided by the number of steps captured) and records them under <key> (e.g. "C2:dense_search"); bench.py reads
the file for its roofline fields - it never runs under a profiler itself."""
import csv
import json
import os
import re
import sys

key, path, pattern = sys.argv[1], sys.argv[2], re.compile(sys.argv[3])
steps = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
note = sys.argv[5] if len(sys.argv) > 5 else ""
rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
hdr = rows[0]
iname, imetric, ivalue, iunit, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6,
         "second": 1e3, "s": 1e3, "inst": 1.0, "": 1.0}
tot, per, launches = {}, {}, set()
by_launch = {}
for r in rows[1:]:
    if len(r) != len(hdr) or not pattern.search(r[iname]):
        continue
    by_launch.setdefault(r[iid], {"name": re.sub(r"\(.*", "", r[iname]).split("::")[-1]})[r[imetric]] = float(r[ivalue].replace(",", "")) * scale.get(r[iunit], 1.0)
for lid, d in by_launch.items():
    launches.add(lid)
    dur = d.get("gpu__time_duration.sum", 0.0)
    k = per.setdefault(d["name"], {})
    for m, v in d.items():
        if m == "name":
            continue
        if m.endswith(".sum"):  # additive
            tot[m] = tot.get(m, 0.0) + v
            k[m] = k.get(m, 0.0) + v
        else:                   # percentages / ratios: mean weighted by the launch's duration
            k[m + " (x ms)"] = k.get(m + " (x ms)", 0.0) + v * dur
for k in per.values():
    dur = k.get("gpu__time_duration.sum", 0.0)
    for m in [m for m in k if m.endswith(" (x ms)")]:
        k[m[:-7]] = k.pop(m) / dur if dur > 0 else 0.0
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "kernel_counters.json")
db = json.load(open(out)) if os.path.exists(out) else {}
db[key] = {"inst_executed": tot.get("smsp__inst_executed.sum", 0.0) / steps,
           "dram_bytes": (tot.get("dram__bytes_read.sum", 0.0) + tot.get("dram__bytes_write.sum", 0.0)) / steps,
           "kernel_ms_under_ncu": tot.get("gpu__time_duration.sum", 0.0) / steps, "launches": len(launches) / steps,
           "per_kernel": {k: {m: v / steps for m, v in d.items()} for k, d in per.items()},
           "source": f"{os.path.relpath(path, os.path.dirname(out) + '/..')} ({note})" if note else os.path.relpath(path, os.path.dirname(out) + "/..")}
json.dump(db, open(out, "w"), indent=1, sort_keys=True)
print(key, {k: v for k, v in db[key].items() if k != "per_kernel"})
