"""One line per captured launch of gpurun_out/prof_render.ncu-rep (tools/ncu_render.sh): python tools/ncu_render_table.py [rep] [out]"""
import csv, subprocess, sys
rep = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/prof_render.ncu-rep"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def val(r, k, scale=1.0):
    try:
        return float(r[ix[k]].replace(",", "")) * scale
    except (KeyError, ValueError):
        return float("nan")


def to_bytes(r, k):
    u = units[ix[k]].lower()
    return val(r, k, {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0))


def to_us(r, k):
    u = units[ix[k]].lower()
    return val(r, k, {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0))


lines = ["%-46s %8s %9s %9s %9s %10s %9s %8s %7s %9s %8s %11s" % ("kernel", "us", "rdMB", "wrMB", "dram%", "inst", "issue%", "alu%",
                                                                   "occ%", "regs", "grid", "DRAM GB/s")]
for r in rows[2:]:
    name = r[ix["Kernel Name"]].replace("(anonymous namespace)::", "").replace("void ", "")
    name = name.split("(")[0][:46]
    us = to_us(r, "gpu__time_duration.sum")
    rd, wr = to_bytes(r, "dram__bytes_read.sum"), to_bytes(r, "dram__bytes_write.sum")
    lines.append("%-46s %8.1f %9.1f %9.1f %9.1f %10.2e %9.1f %8.1f %7.1f %9.0f %8.0f %11.0f" % (
        name, us, rd / 1e6, wr / 1e6, val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        val(r, "smsp__inst_executed.sum"), val(r, "sm__inst_issued.avg.pct_of_peak_sustained_active")
        if "sm__inst_issued.avg.pct_of_peak_sustained_active" in ix else val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        val(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), val(r, "launch__registers_per_thread"),
        val(r, "launch__grid_size"), (rd + wr) / us / 1e3))
txt = "\n".join(lines)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
