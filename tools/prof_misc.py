"""Runs the HBM-bound kernels once each at representative sizes (for ncu)."""
import sys
import torch
sys.path.insert(0, ".")
from human_instance_segmentation_b200 import engine, lib as L

p = engine.Plan(torch.device("cuda")); lib = p.lib
st = torch.cuda.current_stream().cuda_stream
g = torch.Generator().manual_seed(0)
N = 16
def dw(h, w, c, k, s):
    x = p.act(N, h, w, c); x.buf.normal_()
    pad = ((s - 1) + (k - 1)) // 2
    ho, wo = (h + 2 * pad - k) // s + 1, (w + 2 * pad - k) // s + 1
    out = p.act(N, ho, wo, c)
    parts = lib.his_depthwise_pool_parts(N, h, w, c, k, s)
    pool = p.f32(N, parts, c)
    wdw = p.const(torch.randn(k * k, c), torch.float16)
    sc, sh = p.const(torch.ones(c)), p.const(torch.zeros(c))
    L.check(lib.his_depthwise_conv(x.ptr, N, h, w, c, x.cs, wdw.data_ptr(), sc.data_ptr(), sh.data_ptr(), k, s, 2, out.ptr, out.cs, pool.data_ptr(), 0, st))
    gate = p.f32(N, c); gate.fill_(0.5)
    L.check(lib.his_scale_channels(out.ptr, out.cs, gate.data_ptr(), N, ho * wo, c, out.ptr, out.cs, 0, st))
dw(240, 320, 96, 3, 2)
dw(120, 160, 144, 3, 1)
dw(60, 80, 240, 5, 1)
# head-side glue at 160 ROIs
n = 160
x = p.act(n, 64, 48, 256); x.buf.normal_()
sa = p.act(n, 64, 48, 256); stats = p.f32(n, 64, 48, 2)
L.check(lib.his_spatial_attention(x.ptr, n, 64, 48, 256, x.cs, p.const(torch.randn(2, 7, 7)).data_ptr(), 7, stats.data_ptr(), sa.ptr, sa.cs, 0, st))
nchw = p.f32(n, 256, 64, 48)
L.check(lib.his_nhwc_half_to_nchw_float(x.ptr, n, 64 * 48, 256, x.cs, nchw.data_ptr(), 0, st))
torch.cuda.synchronize()
print("ok")
