"""Short per-kernel summary of an ncu report:  python profiles/ncu_extract.py <report.ncu-rep> [<title>]
(duration, DRAM bytes and rate, issue-slot / L1 / L2 / DRAM throughput, occupancy, top stall reasons per issued instruction)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
print(title)
for r in rows[2:]:
    name = r[col["Kernel Name"]].replace("<unnamed>::", "").split("(")[0].replace("void ", "")
    print(f"\n== {name}  (launch id {r[col['ID']]})")
    for w in want:
        if w in col:
            print(f"   {w:62s} {r[col[w]]:>16s} {unit[col[w]]}")
    try:
        t = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
        tu = unit[col["gpu__time_duration.sum"]]
        t_s = t * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(tu, 1e-6)
        def gb(k):
            v = float(r[col[k]].replace(",", ""))
            return v * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(unit[col[k]], 1e-9)
        tot = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
        print(f"   {'DRAM traffic / duration':62s} {tot / t_s / 1e3:16.3f} TB/s  ({tot:.3f} GB)")
    except Exception:
        pass
    top = sorted(((float(r[col[h]].replace(",", "") or 0), h) for h in stall if r[col[h]]), reverse=True)[:4]
    for v, h in top:
        print(f"   stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:56s} {v:16.2f} per issue")
