"""Sums dram bytes of the launches in an `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`
log and prints {kernel substring: bytes per launch}.  usage: python tools/ncu_traffic.py log.csv substring [substring...]"""
import csv, json, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
i_name, i_metric, i_unit, i_val, i_id = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value"), hdr.index("ID")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
out = {}
for sub in sys.argv[2:]:
    ids, total, t_ns = set(), 0.0, 0.0
    for r in rows[1:]:
        if sub not in r[i_name]:
            continue
        ids.add(r[i_id])
        v = float(r[i_val].replace(",", ""))
        if r[i_metric].startswith("dram__bytes"):
            total += v * scale.get(r[i_unit], 1.0)
        elif r[i_metric].startswith("gpu__time_duration"):
            t_ns += v * {"ns": 1.0, "us": 1e3, "usecond": 1e3, "nsecond": 1.0, "ms": 1e6, "msecond": 1e6}.get(r[i_unit], 1.0)
    out[sub] = {"launches": len(ids), "dram_bytes_total": total, "dram_bytes_per_launch": total / max(len(ids), 1), "time_ms_total": t_ns / 1e6}
print(json.dumps(out, indent=1))
