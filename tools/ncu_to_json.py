"""ncu raw CSV of tools/ncu_forward.py (one forward) + gpurun_out/ncu_forward_layers.json -> profiles/r2_ncu_summary.json,
the file bench.py reads at run time for `roofline.traffic` and `roofline.tensor_pipe_util`.
usage: ncu -i rep.ncu-rep --page raw --csv > raw.csv; python tools/ncu_to_json.py raw.csv layers.json out.json [source-name]"""
import csv, json, sys
raw, layers, out = sys.argv[1:4]
source = sys.argv[4] if len(sys.argv) > 4 else raw
rows = list(csv.reader(open(raw)))
hdr = rows[0]
data = [dict(zip(hdr, r)) for r in rows[2:]]
lay = json.load(open(layers))
B = lay["batch"]


def num(d, k):
    try:
        return float(d[k].replace(",", ""))
    except (KeyError, ValueError):
        return None


def to_bytes(d, k, units):
    v = num(d, k)
    if v is None:
        return None
    u = units.get(k, "byte").lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


units = dict(zip(hdr, rows[1]))
# the conv launches in order: everything except the cast / nonlinearity helpers
convs = [d for d in data if any(t in d.get("Kernel Name", "") for t in ("k_conv", "k_level", "k_head", "k_tail"))]
names = [r["name"] for r in lay["rows"]]
if len(convs) != len(names):
    print("warning: %d conv launches in the capture, %d layers in the forward" % (len(convs), len(names)), file=sys.stderr)
launches = {}
tens, memt, dur = [], [], []
for name, d in zip(names, convs):
    t_us = num(d, "gpu__time_duration.sum")
    tu = units.get("gpu__time_duration.sum", "us")
    t_us = t_us * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(tu, 1)
    tp = num(d, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")
    mt = num(d, "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")
    rd, wr = to_bytes(d, "dram__bytes_read.sum", units), to_bytes(d, "dram__bytes_write.sum", units)
    launches["%s@b%d" % (name, B)] = {
        "kernel": d.get("Kernel Name", "")[:80], "us": t_us, "dram_bytes": (rd or 0) + (wr or 0),
        "tensor_pipe_pct": tp, "tensor_mem_pct": mt,
        "issue_pct": num(d, "sm__inst_issued.avg.pct_of_peak_sustained_active"),
        "dram_pct": num(d, "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
        "registers": num(d, "launch__registers_per_thread"), "grid": num(d, "launch__grid_size"),
    }
    if tp is not None and ("k_conv" in d.get("Kernel Name", "") or "k_level" in d.get("Kernel Name", "")):
        tens.append(tp * t_us); memt.append((mt or 0) * t_us); dur.append(t_us)
summary = {
    "source": source, "batch": B,
    # time-weighted over the tcgen05 launches of the forward, % of elapsed cycles
    "tensor_pipe_util": {"sm__pipe_tensor_cycles_active_pct": sum(tens) / sum(dur) if dur else None,
                         "sm__mem_tensor_cycles_active_pct": sum(memt) / sum(dur) if dur else None,
                         "max_launch_pct": max((v["tensor_pipe_pct"] or 0) for v in launches.values()) if launches else None,
                         "note": "ncu --set full, cold caches, serialised launches; time-weighted mean over the tcgen05 launches"},
    "launches": launches,
}
json.dump(summary, open(out, "w"), indent=1)
print("wrote", out, "tensor pipe %.1f %% (time-weighted)" % (summary["tensor_pipe_util"]["sm__pipe_tensor_cycles_active_pct"] or -1))
