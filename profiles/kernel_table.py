"""Reads an `ncu --set full` report (several kernels) and prints one markdown row per kernel: duration, DRAM bytes moved, DRAM
throughput, L2 hit rate, warp instructions, threads per instruction, issue-slot use, registers and the top stall reason.
usage: python profiles/kernel_table.py <report.ncu-rep> [--json out.json]"""
import csv
import io
import json
import re
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
SCALE = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12, "ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "s": 1e3, "second": 1e3}


def val(r, name, scaled=False):
    v = r[col[name]].replace(",", "")
    try:
        v = float(v)
    except ValueError:
        return float("nan")
    return v * SCALE.get(units[col[name]].lower(), 1) if scaled else v


stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
table = []
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    m = re.search(r"<(\w+(?:<[^>]*>)?)", name)
    short = m.group(1) if m else name.split("(")[0]
    ms = val(r, "gpu__time_duration.sum", True)
    rd, wr = val(r, "dram__bytes_read.sum", True), val(r, "dram__bytes_write.sum", True)
    top = max(stalls, key=lambda h: val(r, h))
    table.append(dict(kernel=short, ms=ms, dram_read=rd, dram_write=wr, dram_gbs=(rd + wr) / ms / 1e6,
                      dram_pct=val(r, "dram__bytes_read.sum.pct_of_peak_sustained_elapsed") + val(r, "dram__bytes_write.sum.pct_of_peak_sustained_elapsed"), l2_hit=val(r, "lts__t_sector_hit_rate.pct"),
                      warp_inst=val(r, "smsp__inst_executed.sum"), threads_per_inst=val(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
                      issue_pct=val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), regs=val(r, "launch__registers_per_thread"),
                      warps_active_pct=val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                      top_stall=top[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], top_stall_ratio=val(r, top)))
print("| kernel | ms (ncu) | DRAM MB (rd+wr) | DRAM GB/s | DRAM % | L2 hit % | warp inst (M) | thr/inst | issue % | regs | top stall |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for t in table:
    print("| %s | %.3f | %.0f + %.0f | %.0f | %.1f | %.1f | %.1f | %.1f | %.1f | %d | %s %.1f |" % (
        t["kernel"], t["ms"], t["dram_read"] / 1e6, t["dram_write"] / 1e6, t["dram_gbs"], t["dram_pct"], t["l2_hit"], t["warp_inst"] / 1e6,
        t["threads_per_inst"], t["issue_pct"], t["regs"], t["top_stall"], t["top_stall_ratio"]))
if "--json" in sys.argv:
    json.dump(table, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
