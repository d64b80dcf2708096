"""Top stalled SASS instructions of one kernel launch from an exported ncu source page (CSV).
usage: python tools/ncu_top.py file.csv [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
i_src, i_samp, i_exec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = rows[2:]
tot = sum(int(r[i_samp] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {}
for r in data:
    for i, h in stall_cols:
        agg[h] = agg.get(h, 0) + int(r[i] or 0)
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
order = sorted(range(len(data)), key=lambda k: -int(data[k][i_samp] or 0))[:n]
for k in sorted(order):
    r = data[k]
    st = sorted(((int(r[i] or 0), h) for i, h in stall_cols), reverse=True)[:2]
    print(f"{k:5d} {int(r[i_samp]):6d} {100*int(r[i_samp])/max(tot,1):5.1f}%  exec={r[i_exec]:>8s}  {r[i_src].strip()[:70]:70s} {st}")
