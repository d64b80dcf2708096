"""Event-timed micro-benchmark of the implicit-GEMM conv on the shapes of the B0 step (not a bench value; a tuning aid).

    python tools/bench_gemm.py [--stages 4,12,48] [--reps 10]
Prints ms, algorithmic TFLOP/s and the in+out(+res) GB/s of every shape for every ring depth (HIS_GEMM_STAGES cap).
"""
import argparse
import ctypes
import os
import sys

import torch

sys.path.insert(0, ".")
from human_instance_segmentation_b200 import engine  # noqa: E402
from human_instance_segmentation_b200.engine import RES_ADD, RES_MUL, RES_NONE  # noqa: E402

SHAPES = [
    # (label, n, h, w, cin, cout, k, res_mode, transposed)
    ("unet dec4 conv2 16->16 k3 480x640", 16, 480, 640, 16, 16, 3, RES_NONE, False),
    ("unet dec4 conv1 32->16 k3 480x640", 16, 480, 640, 32, 16, 3, RES_NONE, False),
    ("unet dec3 conv1 96->32 k3 240x320", 32, 240, 320, 96, 32, 3, RES_NONE, False),
    ("unet dec3 conv2 32->32 k3 240x320", 32, 240, 320, 32, 32, 3, RES_NONE, False),
    ("unet tail1 packed 32->64 k3 240x320", 64, 240, 320, 32, 64, 3, RES_NONE, False),
    ("unet tail2 packed 64->64 k3 240x320", 64, 240, 320, 64, 64, 3, RES_NONE, False),
    ("unet probe 32->32 k3 240x320 n64", 64, 240, 320, 32, 32, 3, RES_NONE, False),
    ("unet dec2 conv1 152->64 k3 120x160", 64, 120, 160, 152, 64, 3, RES_NONE, False),
    ("unet expand 16->96 k1 240x320", 32, 240, 320, 16, 96, 1, RES_NONE, False),
    ("unet expand 24->144 k1 120x160", 64, 120, 160, 24, 144, 1, RES_NONE, False),
    ("head 64->64 k3 64x48", 640, 64, 48, 64, 64, 3, RES_ADD, False),
    ("head 128->128 k3 64x48", 640, 64, 48, 128, 128, 3, RES_NONE, False),
    ("head 128->128 k3 128x96", 320, 128, 96, 128, 128, 3, RES_NONE, False),
    ("head 256->128 k3 64x48", 640, 64, 48, 256, 128, 3, RES_NONE, False),
    ("head 256->64 k3 64x48", 640, 64, 48, 256, 64, 3, RES_NONE, False),
    ("head 256->256 k3 64x48", 640, 64, 48, 256, 256, 3, RES_NONE, False),
    ("head 256->256 k3 64x48 res", 640, 64, 48, 256, 256, 3, RES_ADD, False),
    ("head 128->256 k1 gate*shared", 640, 64, 48, 128, 256, 1, RES_MUL, False),
    ("head 258->256 k1", 640, 64, 48, 258, 256, 1, RES_NONE, False),
    ("head convT 256->128", 640, 64, 48, 256, 128, 1, RES_NONE, True),
    ("head 256->256 k1 plain", 640, 64, 48, 256, 256, 1, RES_NONE, False),
    ("head 128->128 k3 128x96 res", 320, 128, 96, 128, 128, 3, RES_ADD, False),
    ("head 128->128 k3 64x48 res", 640, 64, 48, 128, 128, 3, RES_ADD, False),
]


def bench(shape, reps, act=1):
    label, n, h, w, cin, cout, k, res_mode, transposed = shape
    dev = torch.device("cuda")
    plan = engine.Plan(dev)
    x = plan.act(n, h, w, cin)
    x.buf.normal_()
    wt = torch.randn(cin, cout, 2, 2) if transposed else torch.randn(cout, cin, k, k)
    nt, bn = ctypes.c_int(), ctypes.c_int()
    plan.lib.his_conv_gemm_tile_n(cout, ctypes.byref(nt), ctypes.byref(bn))
    slab = nt.value * bn.value
    wp, cin_pad = engine.pack_gemm_weight(wt * 0.05, slab, transposed)
    oh, ow = (2 * h, 2 * w) if transposed else (h, w)
    out = plan.act(n, oh, ow, cout)
    res = None
    if res_mode != RES_NONE:
        res = plan.act(n, oh, ow, cout)
        res.buf.normal_()
    plan.conv_gemm(x, plan.const(wp, torch.float16), cin_pad, plan.const(torch.zeros(slab)), out, k, act, 1.0, res,
                   res_mode, transposed)
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
    byts = 2 * (x.buf.numel() + out.buf.numel() + (res.buf.numel() if res is not None else 0))
    return ms, plan.flops / ms / 1e9, byts / ms / 1e6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stages", default="0")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--only", default="")
    ap.add_argument("--act", type=int, default=1, help="0 none, 1 relu, 2 silu, 3 sigmoid")
    args = ap.parse_args()
    sweeps = [int(v) for v in args.stages.split(",")]
    for shape in SHAPES:
        if args.only and args.only not in shape[0]:
            continue
        cells = []
        for st in sweeps:
            if st:
                os.environ["HIS_GEMM_STAGES"] = str(st)
            else:
                os.environ.pop("HIS_GEMM_STAGES", None)
            ms, tf, gbs = bench(shape, args.reps, args.act)
            cells.append(f"st{st}: {ms:7.3f} ms {tf:7.1f} TF {gbs:6.0f} GB/s")
        print(f"{shape[0]:38s} " + " | ".join(cells), flush=True)


if __name__ == "__main__":
    main()
