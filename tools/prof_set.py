"""One launch each of the kernels under study (for `ncu --set full`): depthwise variants, narrow / medium / hot GEMMs,
the GEMMs with the fused fp32 export.  python tools/prof_set.py"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from human_instance_segmentation_b200 import engine, lib as L  # noqa: E402
from human_instance_segmentation_b200.engine import RES_ADD, RES_MUL, RES_NONE  # noqa: E402

dev = torch.device("cuda")
p = engine.Plan(dev)
lib = p.lib
st = torch.cuda.current_stream().cuda_stream


def dw(N, h, w, c, k, s):
    x = p.act(N, h, w, c); x.buf.normal_()
    pad = ((s - 1) + (k - 1)) // 2
    ho, wo = (h + 2 * pad - k) // s + 1, (w + 2 * pad - k) // s + 1
    out = p.act(N, ho, wo, c)
    parts = lib.his_depthwise_pool_parts(N, h, w, c, k, s)
    pool = p.f32(N, parts, c)
    wdw = p.const(torch.randn(k * k, c), torch.float16)
    sc, sh = p.const(torch.ones(c)), p.const(torch.zeros(c))
    L.check(lib.his_depthwise_conv(x.ptr, N, h, w, c, x.cs, wdw.data_ptr(), sc.data_ptr(), sh.data_ptr(), k, s, 2, out.ptr, out.cs, pool.data_ptr(), 0, st))


def gemm(n, h, w, cin, cout, k, res_mode=RES_NONE, aux=False, act=1):
    x = p.act(n, h, w, cin); x.buf.normal_()
    wt = torch.randn(cout, cin, k, k) * 0.05
    nt, bn = ctypes.c_int(), ctypes.c_int()
    lib.his_conv_gemm_tile_n(cout, ctypes.byref(nt), ctypes.byref(bn))
    slab = nt.value * bn.value
    wp, cin_pad = engine.pack_gemm_weight(wt, slab, False)
    out = p.act(n, h, w, cout)
    res = None
    if res_mode != RES_NONE:
        res = p.act(n, h, w, cout); res.buf.normal_()
    auxt = p.f32(n, cout, h, w) if aux else None
    p.conv_gemm(x, p.const(wp, torch.float16), cin_pad, p.const(torch.zeros(slab)), out, k, act, 1.0, res, res_mode,
                False, aux_f32=auxt)


dw(64, 60, 80, 240, 5, 1)
dw(64, 120, 160, 144, 5, 2)
dw(64, 120, 160, 144, 3, 1)
dw(64, 240, 320, 96, 3, 2)
gemm(16, 480, 640, 16, 16, 3)
gemm(32, 240, 320, 96, 32, 3)
gemm(640, 64, 48, 64, 64, 3, RES_ADD)
gemm(320, 128, 96, 128, 128, 3)
gemm(640, 64, 48, 128, 256, 1, RES_MUL, aux=True, act=3)
gemm(640, 64, 48, 256, 256, 3, RES_ADD, aux=True)
gemm(640, 64, 48, 256, 256, 3, RES_ADD)
p.replay()
# post-processing: the fused clean-up chain on 64 full-image masks (BASELINE config 5)
from human_instance_segmentation_b200 import postprocess as pp  # noqa: E402
masks = (torch.rand(64, 1, 480, 640, device=dev) > 0.5).float()
pp.MaskCleanup().to(dev)(masks)
torch.cuda.synchronize()
print("ok")
