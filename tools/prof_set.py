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


def gemm(n, h, w, cin, cout, k, res_mode=RES_NONE, aux=False, act=1, transposed=False, plan=None):
    q = plan or p
    x = q.act(n, h, w, cin); x.buf.normal_()
    wt = (torch.randn(cin, cout, 2, 2) if transposed else torch.randn(cout, cin, k, k)) * 0.05
    nt, bn = ctypes.c_int(), ctypes.c_int()
    lib.his_conv_gemm_tile_n(cout, ctypes.byref(nt), ctypes.byref(bn))
    slab = nt.value * bn.value
    wp, cin_pad = engine.pack_gemm_weight(wt, slab, transposed, split=q.split)
    oh, ow = (2 * h, 2 * w) if transposed else (h, w)
    out = q.act(n, oh, ow, cout)
    res = None
    if res_mode != RES_NONE:
        res = q.act(n, oh, ow, cout); res.buf.normal_()
    auxt = q.f32(n, cout, h, w) if aux else None
    q.conv_gemm(x, q.const(wp, torch.float16), cin_pad, q.const(torch.zeros(slab)), out, k, act, 1.0, res, res_mode,
                transposed, aux_f32=auxt)


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
gemm(640, 64, 48, 256, 128, 1, transposed=True)            # ConvTranspose k2s2, two merged phases per N tile
gemm(32, 240, 320, 16, 96, 1, act=2)                       # MBConv expand + SiLU, three epilogue groups
p.replay()
# strict precision: the split-fp16 kernel on the hot shape (three MMA passes, hi / lo planes)
ps = engine.Plan(dev, True)
gemm(640, 64, 48, 256, 256, 3, RES_ADD, plan=ps)
ps.replay()
# both RoI aligners of the B0 step in one launch (warp per ROI row)
sys.path.insert(0, ".")
from human_instance_segmentation_b200.synthetic import synth_images, synth_rois  # noqa: E402
imgs, rois = synth_images(1, 64, 480, 640).to(dev), synth_rois(1, 64, 10).to(dev)
two = torch.randn(64, 2, 480, 640, device=dev)
comb, patches = p.act(640, 64, 48, 264), p.act_zeroed(640, 64, 48, 3)
rf, rp = p.f32(640, 2, 64, 48), p.f32(640, 3, 64, 48)
msk = comb.slice(256, 2)
L.check(lib.his_roi_align_fused(two.data_ptr(), 2, 480.0, 640.0, 1, msk.ptr, msk.cs, rf.data_ptr(), imgs.data_ptr(), 3, 480.0, 640.0, 1, patches.ptr,
                                patches.cs, rp.data_ptr(), 64, 480, 640, rois.data_ptr(), 640, 64, 48, 0, st))
# post-processing: the fused clean-up chain on 64 full-image masks (BASELINE config 5)
from human_instance_segmentation_b200 import postprocess as pp  # noqa: E402
masks = (torch.rand(64, 1, 480, 640, device=dev) > 0.5).float()
pp.MaskCleanup().to(dev)(masks)
# ROI chain of config 5: dilation + argmax in one pass, paste-back (640 ROIs of 64 images)
lg = torch.nn.functional.interpolate(torch.randn(640, 3, 16, 12, device=dev) * 2, size=(128, 96), mode="bilinear").contiguous()
pp.paste_masks(pp.instance_masks(lg, as_uint8=True, dilation_pixels=1), rois, 64, 480, 640)
torch.cuda.synchronize()
print("ok")
