"""Per-kernel table of an `ncu --csv --metrics ...` launch list (every launch of a command): launches, total and mean
time, DRAM bytes, L2->L1 bytes, issue-slot and tensor-pipe utilisation.   usage: python tools/launch_table.py launches.csv"""
import csv, collections, re, sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[ix["ID"]] == "ID":
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
    key = (r[ix["ID"]], name)
    agg.setdefault(key, {})[r[ix["Metric Name"]]] = (float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]])
per = collections.OrderedDict()
for (_, name), m in agg.items():
    e = per.setdefault(name, collections.defaultdict(float))
    e["n"] += 1
    t, u = m.get("gpu__time_duration.sum", (0, "ns"))
    e["ms"] += t * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    def byts(k):
        v, u = m.get(k, (0, "byte"))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
    e["dram"] += byts("dram__bytes_read.sum") + byts("dram__bytes_write.sum")
    e["l2l1"] += byts("l1tex__m_xbar2l1tex_read_bytes.sum")
    tt = t * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    e["issue_w"] += m.get("smsp__issue_active.avg.pct_of_peak_sustained_active", (0, ""))[0] * tt
    e["tensor_w"] += m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", (0, ""))[0] * tt
tot = sum(e["ms"] for e in per.values())
print(f"{'kernel':58s} {'n':>5s} {'ms':>10s} {'share':>6s} {'DRAM GB':>9s} {'L2->L1 GB':>10s} {'issue%':>7s} {'tensor%':>8s}")
for name, e in sorted(per.items(), key=lambda kv: -kv[1]["ms"]):
    ms = e["ms"]
    print(f"{name[:58]:58s} {int(e['n']):5d} {ms:10.3f} {100 * ms / tot:5.1f}% {e['dram'] / 1e9:9.3f} {e['l2l1'] / 1e9:10.2f} "
          f"{e['issue_w'] / ms if ms else 0:7.1f} {e['tensor_w'] / ms if ms else 0:8.1f}")
print(f"total {tot:.3f} ms under ncu (cold caches, serialised: shares, not absolutes)")
