"""Experiment: why are Cin=72 / 144 / 152 layers slow?  Varies the buffer stride (cs) and the TMA channel extent independently."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from human_instance_segmentation_b200 import engine  # noqa: E402
from human_instance_segmentation_b200.engine import RES_NONE  # noqa: E402


def bench(n, h, w, cin, cout, k, in_total=None, out_total=None, reps=10):
    dev = torch.device("cuda")
    plan = engine.Plan(dev)
    xb = plan.act(n, h, w, in_total or cin)
    xb.buf.normal_()
    x = xb.slice(0, cin)
    wt = torch.randn(cout, cin, k, k) * 0.05
    nt, bn = ctypes.c_int(), ctypes.c_int()
    plan.lib.his_conv_gemm_tile_n(cout, ctypes.byref(nt), ctypes.byref(bn))
    slab = nt.value * bn.value
    wp, cin_pad = engine.pack_gemm_weight(wt, slab, False)
    ob = plan.act(n, h, w, out_total or cout)
    out = ob.slice(0, cout)
    plan.conv_gemm(x, plan.const(wp, torch.float16), cin_pad, plan.const(torch.zeros(slab)), out, k, 1, 1.0, None, RES_NONE, False)
    for _ in range(2):
        plan.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        plan.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms, plan.flops / ms / 1e9


for (cin, cout, it, ot) in [(72, 72, None, None), (72, 72, 96, None), (72, 72, 128, None), (72, 72, None, 96), (72, 72, 96, 96), (96, 72, None, None), (64, 72, None, None),
                            (80, 72, None, None), (96, 96, None, None), (72, 96, None, None), (72, 128, None, None), (64, 64, None, None),
                            (144, 72, None, None), (144, 72, 160, None), (160, 72, None, None), (128, 72, None, None)]:
    ms, tf = bench(416, 80, 60, cin, cout, 3, it, ot)
    print(f"cin {cin:4d} (cs {it or cin:4d}) cout {cout:4d} (cs {ot or cout:4d}): {ms:7.3f} ms {tf:7.1f} TF", flush=True)
