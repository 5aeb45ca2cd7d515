"""Summaries of round-2 gpurun outputs: python profiles/summarize_r2.py bench <json...> | launches <csv> [iters]"""
import csv, io, json, re, sys

def bench(files):
    for fn in files:
        try:
            d = json.load(open(fn))
        except Exception as ex:
            print(fn, "unreadable", ex); continue
        ph = {k: round(v, 2) for k, v in d.get("phase_ms_per_step_rank0", {}).items()}
        r = d.get("roofline", {})
        print(f"{fn}: loci/s={d['value']:.0f} ms/step={d['ms_per_step']:.2f} e2e_ms={d['e2e']['ms_per_step']:.1f} {ph} "
              f"frac={r.get('frac')} share={r.get('kernel_share_of_step')} clocks={d.get('clocks')}")

def launches(fn, iters=3, min_us=10.0):
    lines = open(fn).read().splitlines()
    i = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[i:]))))
    agg, order = {}, []
    for r in rows:
        k = r["ID"]
        if k not in agg:
            agg[k] = {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("<unnamed>::", "").replace("void ", "")}
            order.append(k)
        agg[k][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    per = len(order) // iters
    tot = 0.0
    for k in order[(iters - 1) * per:]:
        a = agg[k]
        t = a["gpu__time_duration.sum"] / 1e3
        tot += t
        rd, wr = a.get("dram__bytes_read.sum", 0) / 1e6, a.get("dram__bytes_write.sum", 0) / 1e6
        if t >= min_us:
            print(f"{a['name'][:44]:44s} {t:9.1f} us  rd {rd:9.1f} MB  wr {wr:9.1f} MB  {(rd + wr) / t / 1e3:6.3f} TB/s")
    print(f"sum of the last of {iters} iterations ({per} launches): {tot / 1e3:.3f} ms")

if __name__ == "__main__":
    if sys.argv[1] == "bench":
        bench(sys.argv[2:])
    else:
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 3)
