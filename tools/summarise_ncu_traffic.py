"""Summarise an `ncu --csv --page raw` launch list (metrics gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum)
into (a) per-kernel totals/shares and (b) profiles/r2_conv_dram_traffic.json used by bench.py's roofline.traffic.
Usage: python tools/summarise_ncu_traffic.py <launches.csv> [out.json]"""
import csv, json, re, sys, collections

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
long_form = "Metric Name" in hdr
if long_form:  # one row per (launch, metric)
    ix = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict()
    for r in rd:
        if len(r) < len(hdr):
            continue
        d = per.setdefault(r[ix["ID"]], {"name": r[ix["Kernel Name"]]})
        v = float(r[ix["Metric Value"]].replace(",", ""))
        u = r[ix["Metric Unit"]]
        m = r[ix["Metric Name"]]
        if m.startswith("gpu__time_duration"):
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
        elif "bytes" in m:
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        d[m] = v
    rows = list(per.values())
else:
    units = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rd:
        d = {"name": r[ix["Kernel Name"]]}
        for m in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"):
            if m in ix:
                v = float(r[ix[m]].replace(",", ""))
                u = units[ix[m]]
                if m.startswith("gpu__time"):
                    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(u, 1e-6)
                else:
                    v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                d[m] = v
        rows.append(d)

agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for d in rows:
    k = re.sub(r"\(.*", "", d["name"])
    k = re.sub(r"^void ", "", k)
    a = agg[k]
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
print("launches %d  total %.2f ms (serialised, cold-cache)" % (len(rows), tot))
print("%-70s %7s %10s %7s %12s" % ("kernel", "count", "ms", "share", "dram GB"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%-70s %7d %10.3f %6.1f%% %12.3f" % (k[:70], a[0], a[1], 100 * a[1] / tot, a[2] / 1e9))
conv = [a for k, a in agg.items() if "conv_igemm_kernel" in k or "conv_pixn_kernel" in k or "conv_pair_kernel" in k]
if conv and len(sys.argv) > 2:
    n = sum(a[0] for a in conv); b = sum(a[2] for a in conv); ms = sum(a[1] for a in conv)
    if b > 0:
        json.dump({"kernel": "conv_pair_kernel + conv_pixn_kernel + conv_igemm_kernel", "launches": n, "avg_dram_bytes_per_launch": b / n, "total_dram_bytes": b,
                   "total_ms_under_ncu": ms, "source": sys.argv[1]}, open(sys.argv[2], "w"), indent=1)
        print("wrote", sys.argv[2])
