"""Reads an `ncu --set full` report and records the DRAM traffic of the captured kernel in profiles/kernel_traffic.json
(the `roofline.traffic` field of bench.py).  usage: python profiles/extract_traffic.py <report.ncu-rep> <key> [kernel-name substring]
where key is e.g. "InsertBody<2>:cfg2:k51"; with a substring the first launch whose name contains it is taken (reports with several kernels)."""
import csv
import io
import json
import os
import subprocess
import sys

rep, key = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
want = sys.argv[3] if len(sys.argv) > 3 else ""
vals = next(r for r in rows[2:] if want in r[col["Kernel Name"]])


def to_bytes(name):
    v, u = float(vals[col[name]].replace(",", "")), units[col[name]].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}[u]


rec = {"dram_bytes": int(to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")),
       "dram_read": int(to_bytes("dram__bytes_read.sum")), "dram_write": int(to_bytes("dram__bytes_write.sum")),
       "duration_ms_under_ncu": float(vals[col["gpu__time_duration.sum"]].replace(",", "")) * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}[units[col["gpu__time_duration.sum"]]],
       "report": os.path.basename(rep), "kernel": vals[col["Kernel Name"]][:120]}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kernel_traffic.json")
data = json.load(open(path)) if os.path.exists(path) else {}
data[key] = rec
json.dump(data, open(path, "w"), indent=1, sort_keys=True)
print(key, rec)
