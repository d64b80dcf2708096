"""Copies the measurement pass of tools/gpu_final.sh from gpurun_out/ into profiles/ (tracked) under round-2 names and derives
profiles/r2_traffic.json (DRAM bytes per launch of the dominant kernels, read by bench.py for `roofline.traffic`).
usage: python tools/collect_profiles.py [prefix=r2f]"""
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = sys.argv[1] if len(sys.argv) > 1 else "r2f"
SRC, DST = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
NAMES = ["bench_b0.json", "bench_b0_breakdown.txt", "bench_ref.json", "bench_b1.json", "bench_b1_breakdown.txt", "bench_b7.json",
         "bench_b7_breakdown.txt", "bench_b0_160x120.json", "bench_b0_160x120_breakdown.txt", "bench_strict.json",
         "bench_strict_breakdown.txt", "bench_post.json", "launches.csv", "post_launches.csv", "prof_set_raw.csv", "tests.log"]
for n in NAMES:
    src = os.path.join(SRC, f"{P}_{n}")
    if os.path.exists(src):
        shutil.copy(src, os.path.join(DST, f"r2_{n}"))


def traffic(path, sub):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    i_name, i_metric, i_unit, i_val, i_id = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    ids, total = set(), 0.0
    for r in rows[1:]:
        if sub in r[i_name] and r[i_metric].startswith("dram__bytes"):
            ids.add(r[i_id])
            total += float(r[i_val].replace(",", "")) * scale.get(r[i_unit], 1.0)
    return total / max(len(ids), 1)


out = {"_source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum of one step (profiles/r2_launches.csv, r2_post_launches.csv), bytes per launch"}
lp, pp = os.path.join(DST, "r2_launches.csv"), os.path.join(DST, "r2_post_launches.csv")
if os.path.exists(lp):
    out["conv_gemm_sm100_kernel:b0"] = traffic(lp, "conv_gemm")
    out["depthwise:b0"] = traffic(lp, "depthwise")
    out["roi_align_rows_kernel:b0"] = traffic(lp, "roi_align")
if os.path.exists(pp):
    out["mask_cleanup_wide_kernel:post"] = traffic(pp, "mask_cleanup")
json.dump(out, open(os.path.join(DST, "r2_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
