"""Summarise gpurun_out/resample_{details.txt,raw.csv,source.csv} (tools/ncu_resample.sh): headline metrics, stall
reasons per issued instruction, and stall samples per 250-instruction region of the SASS."""
import csv, os, re, sys
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
for line in open(os.path.join(root, "resample_details.txt")):
    if re.search(r"Duration|Executed Ipc Active|Issue Slots Busy|Registers Per|Theoretical Occ|No Eligible|Mem Busy|Mem Pipes Busy|Warp Cycles Per Issued|Block Limit (Reg|Shared)", line):
        print(line.rstrip())
rows = list(csv.reader(open(os.path.join(root, "resample_raw.csv"))))
d = dict(zip(rows[0], rows[-1]))
for k in d:
    if "smsp__average_warps_issue_stalled" in k and k.endswith("per_issue_active.ratio"):
        try:
            v = float(d[k].replace(",", ""))
        except ValueError:
            continue
        if v > 0.15:
            print("stall", k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), round(v, 2))
for k in ["sm__inst_executed_pipe_alu.sum.pct", "sm__inst_executed_pipe_fma.sum.pct", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
          "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
          "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum"]:
    for kk in d:
        if kk.startswith(k) and "per_second" not in kk:
            print(kk, d[kk])
rows = list(csv.reader(open(os.path.join(root, "resample_source.csv"))))
hdr, data = rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ci["# Samples"]]) for r in data)
nw = max(int(r[ci["Instructions Executed"]]) for r in data[:5])
print("instructions", len(data), "samples", tot, "warps", nw)
seg = int(sys.argv[1]) if len(sys.argv) > 1 else 250
keys = ["stall_short_sb", "stall_wait", "stall_no_inst", "stall_long_sb", "stall_math", "stall_selected", "stall_not_selected",
        "stall_branch_resolving", "stall_mio", "stall_lg", "stall_dispatch"]
for s0 in range(0, len(data), seg):
    rs = data[s0:s0 + seg]
    smp = sum(int(r[ci["# Samples"]]) for r in rs)
    ex = sum(int(r[ci["Instructions Executed"]]) for r in rs)
    st = {k.replace("stall_", ""): sum(int(r[ci[k]]) for r in rs) for k in keys}
    ops = {}
    for r in rs:
        t = r[ci["Source"]].split()
        op = t[1] if t[0].startswith("@") else t[0]
        ops[op] = ops.get(op, 0) + 1
    top = sorted(ops.items(), key=lambda x: -x[1])[:3]
    print("%5d  %5.1f%% of samples  %5.0f instr/warp  %s  %s" % (s0, 100 * smp / max(tot, 1), ex / max(nw, 1),
                                                               {k: v for k, v in st.items() if v > 0.04 * smp and v > 5}, top))
