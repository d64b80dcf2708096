"""Per-source-line stall samples of one kernel launch (ncu --page source --print-source cuda --csv).
usage: python tools/ncu_lines.py report.ncu-rep launch_skip [kernel_regex] [top_n]"""
import csv, io, subprocess, sys
rep, skip = sys.argv[1], sys.argv[2]
rx = sys.argv[3] if len(sys.argv) > 3 else "conv_gemm"
n = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda", "--csv", "--kernel-name", f"regex:{rx}",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1][:150])
hdr = rows[1]
i_src, i_samp = hdr.index("Source"), hdr.index("# Samples")
i_line = hdr.index("Line") if "Line" in hdr else 0
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[2:] if len(r) > i_samp]
tot = sum(int(r[i_samp] or 0) for r in data)
print("total samples", tot)
top = sorted(data, key=lambda r: -int(r[i_samp] or 0))[:n]
keep = set(id(r) for r in top)
for r in data:
    if id(r) in keep and int(r[i_samp] or 0) > 0:
        st = sorted(((int(r[i] or 0), h) for i, h in stall_cols), reverse=True)[:2]
        print(f"{r[i_line]:>5s} {int(r[i_samp]):7d} {100*int(r[i_samp])/max(tot,1):5.1f}%  {r[i_src].strip()[:110]:110s} {st}")
