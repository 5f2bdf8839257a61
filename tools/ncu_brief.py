"""Brief view of an .ncu-rep: key raw metrics + top stalled SASS lines.  usage: ncu_brief.py file.ncu-rep [ntop]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "sm__cycles_active.avg", "smsp__inst_executed.sum", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "smsp__warps_active.avg.per_cycle_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__m_xbar2l1tex_read_bytes.sum"]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print("==", d.get("Kernel Name", "")[:100])
    for k in KEYS:
        if k in d:
            print("  %-70s %s %s" % (k, d[k], units[hdr.index(k)]))
    st = {k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""): float(v) for k, v in d.items()
          if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v}
    print("  stalls/issue:", " ".join("%s=%.2f" % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:9]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("total samples", tot, "sass lines", len(data), "warp-inst", sum(int(r[ix["Instructions Executed"]]) for r in data))
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:ntop]:
    print("%5s %9s lsb=%-4s bar=%-4s ssb=%-4s wait=%-4s math=%-3s %s" % (
        r[ix["# Samples"]], r[ix["Instructions Executed"]], r[ix["stall_long_sb"]], r[ix["stall_barrier"]], r[ix["stall_short_sb"]],
        r[ix["stall_wait"]], r[ix["stall_math"]], r[ix["Source"]].strip()[:80]))
