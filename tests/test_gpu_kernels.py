"""GPU: the non-GEMM kernels through the C-ABI wrappers against golden vectors / plain torch fp32 references."""
import pytest
import torch
import torch.nn.functional as F

import human_instance_segmentation_b200 as his
from human_instance_segmentation_b200 import engine, lib as L
from tests import common

pytestmark = pytest.mark.gpu


def test_dynamic_roi_align_matches_reference_golden():
    g = common.golden("roi_align")
    feat, rois = g["feat"].cuda(), g["rois"].cuda()
    for tag, scale, aligned, (oh, ow) in [("a640", 640.0, True, (16, 12)), ("ahw", (37.0, 53.0), True, (16, 12)),
                                          ("u_hw", (37.0, 53.0), False, (7, 9)), ("a64", 64.0, True, (5, 3))]:
        ra = his.DynamicRoIAlign(spatial_scale=scale, sampling_ratio=2, aligned=aligned)
        out = ra(feat, rois, oh, ow).cpu()
        assert (out - g[tag]).abs().max() < 2e-5, tag
    assert his.DynamicRoIAlign(640.0)(feat, rois[:0], 4, 4).shape == (0, 5, 4, 4)
    # channels-last / non-contiguous feature maps go through the stride arguments
    out = his.DynamicRoIAlign((37.0, 53.0), aligned=True)(feat.to(memory_format=torch.channels_last), rois, 16, 12).cpu()
    assert (out - g["ahw"]).abs().max() < 2e-5


def _plan(split=False):
    return engine.Plan(torch.device("cuda"), split)


SPLIT = pytest.mark.parametrize("split", [False, True], ids=["fp16", "split"])


@SPLIT
@pytest.mark.parametrize("k,s,c", [(3, 1, 32), (3, 2, 96), (5, 2, 144), (5, 1, 672), (3, 1, 1152), (3, 1, 16), (5, 1, 240), (3, 2, 40)])
def test_depthwise_se_matches_torch(k, s, c, split):
    p = _plan(split)
    lib = p.lib
    S = 1 if split else 0
    tol = 2e-5 if split else 2e-3
    g = torch.Generator().manual_seed(k * 100 + c)
    n, h, w = 2, 23, 31
    x = p.act(n, h, w, c); x.fill_nhwc(torch.randn(n, h, w, c, generator=g))
    wt = torch.randn(c, 1, k, k, generator=g) * 0.3
    scale, shift = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.1
    pad = ((s - 1) + (k - 1)) // 2
    ho, wo = (h + 2 * pad - k) // s + 1, (w + 2 * pad - k) // s + 1
    out = p.act(n, ho, wo, c)
    parts = lib.his_depthwise_pool_parts(n, h, w, c, k, s)
    pool = p.f32(n, parts, c)
    wdw = p.const(wt.reshape(c, k * k).t().contiguous(), torch.float32 if split else torch.float16)      # strict mode: fp32 taps
    st = torch.cuda.current_stream().cuda_stream
    L.check(lib.his_depthwise_conv(x.ptr, n, h, w, c, x.cs, wdw.data_ptr(), p.const(scale).data_ptr(), p.const(shift).data_ptr(), k, s, 2,
                                   out.ptr, out.cs, pool.data_ptr(), S, st))
    r = max(c // 4, 1)
    w1, b1 = torch.randn(r, c, generator=g) * 0.2, torch.randn(r, generator=g) * 0.1
    w2, b2 = torch.randn(c, r, generator=g) * 0.2, torch.randn(c, generator=g) * 0.1
    gate = p.f32(n, c)
    L.check(lib.his_se_gate(pool.data_ptr(), parts, n, ho * wo, c, r, p.const(w1).data_ptr(), p.const(b1).data_ptr(), p.const(w2).data_ptr(),
                            p.const(b2).data_ptr(), 2, 1.0, p.f32(n, r).data_ptr(), gate.data_ptr(), st))
    scaled = p.act(n, ho, wo, c)
    L.check(lib.his_scale_channels(out.ptr, out.cs, gate.data_ptr(), n, ho * wo, c, scaled.ptr, scaled.cs, S, st))
    torch.cuda.synchronize()
    ref = F.silu(F.conv2d(x.torch_nchw().cpu(), wt if split else wt.half().float(), None, s, pad, 1, c) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    got = out.torch_nchw().cpu()
    assert (got - ref).abs().max() <= tol * max(1.0, float(ref.abs().max()))
    mean = got.mean((2, 3))
    gref = torch.sigmoid(F.silu(mean @ w1.t() + b1) @ w2.t() + b2)
    assert (gate.cpu() - gref).abs().max() < 1e-4
    assert (scaled.torch_nchw().cpu() - got * gref[:, :, None, None]).abs().max() <= max(tol, 1e-4) * max(1.0, float(ref.abs().max()))


@SPLIT
def test_direct_conv_variants_match_torch(split):
    p = _plan(split)
    wdt = torch.float32 if split else torch.float16
    rnd = (lambda t: t) if split else (lambda t: t.half().float())
    tol = 2e-5 if split else 2e-3
    g = torch.Generator().manual_seed(4)
    # (a) NCHW fp32 input + input affine + stride 2 (the UNet stem)
    n, h, w = 2, 37, 50
    img = torch.rand(n, 3, h, w, generator=g)
    aff = torch.tensor([2.0, 3.0, 4.0, -0.5, 0.25, 0.1])
    wt = torch.randn(32, 3, 3, 3, generator=g) * 0.2
    scale, shift = torch.rand(32, generator=g) + 0.5, torch.randn(32, generator=g) * 0.1
    ho, wo = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
    out = p.act(n, ho, wo, 32)
    p.conv_direct(p.const(img), 1, n, h, w, 3, 0, p.const(engine.pack_direct_weight(wt, f32=split), wdt), p.const(scale), p.const(shift), 32, 3, 2, 1,
                  2, 1.0, in_affine=p.const(aff), out=out)
    # (b) NHWC half input, tail 1x1 to 2 channels, fp32 NCHW output
    x = p.act(n, 9, 11, 128); x.fill_nhwc(torch.randn(n, 9, 11, 128, generator=g))
    w2 = torch.randn(2, 128, 1, 1, generator=g) * 0.1
    tail = p.f32(n, 2, 9, 11)
    p.conv_direct(x, 0, n, 9, 11, 128, x.cs, p.const(engine.pack_direct_weight(w2, f32=split), wdt), p.const(torch.ones(2)), p.const(torch.tensor([0.1, -0.2])),
                  2, 1, 1, 0, 0, out_f32=tail)
    # (c) 3x3 16 -> 1 segmentation head (four output pixels per thread), widths that are / are not multiples of four, input slice
    heads = []
    for hw in ((13, 24), (7, 10), (5, 3)):
        xb = p.act(n, hw[0], hw[1], 24); xb.fill_nhwc(torch.randn(n, hw[0], hw[1], 24, generator=g))
        xs = xb.slice(8, 16)
        w3 = torch.randn(1, 16, 3, 3, generator=g) * 0.1
        o3 = p.f32(n, 1, hw[0], hw[1])
        p.conv_direct(xs, 0, n, hw[0], hw[1], 16, xs.cs, p.const(engine.pack_direct_weight(w3, f32=split), wdt), p.const(torch.tensor([1.5])),
                      p.const(torch.tensor([-0.3])), 1, 3, 1, 1, 3, out_f32=o3)
        heads.append((xs, w3, o3))
    p.replay(); torch.cuda.synchronize()
    for xs, w3, o3 in heads:
        ref3 = torch.sigmoid(F.conv2d(xs.torch_nchw().cpu(), rnd(w3), None, 1, 1) * 1.5 - 0.3)
        assert (o3.cpu() - ref3).abs().max() <= 1e-5
    xin = img * aff[:3].view(1, 3, 1, 1) + aff[3:].view(1, 3, 1, 1)
    ref = F.silu(F.conv2d(xin, rnd(wt), None, 2, 1) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    assert (out.torch_nchw().cpu() - ref).abs().max() <= tol * float(ref.abs().max())
    ref2 = F.conv2d(x.torch_nchw().cpu(), rnd(w2), torch.tensor([0.1, -0.2]))
    assert (tail.cpu() - ref2).abs().max() <= 1e-4 * max(1.0, float(ref2.abs().max()))


@SPLIT
@pytest.mark.parametrize("c,groups", [(72, 1), (72, 9), (64, 4), (256, 256), (32, 8)])
def test_groupnorm_matches_torch(c, groups, split):
    """his_groupnorm_act (GroupNorm / SpatialGroupNorm of get_normalization_layer; one group per channel = instance norm) against
    F.group_norm on the same fp16-rounded input, with residual add + ReLU and with a plain SiLU epilogue."""
    p = _plan(split)
    L = p.lib
    g = torch.Generator().manual_seed(c + groups)
    n, h, w = 3, 13, 11
    x = p.act(n, h, w, c); x.fill_nhwc(torch.randn(n, h, w, c, generator=g) * 1.5 + 0.7)
    r = p.act(n, h, w, c); r.fill_nhwc(torch.randn(n, h, w, c, generator=g))
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.3
    outs = []
    for act, res in ((1, r), (2, None)):
        out = p.act(n, h, w, c)
        parts = L.his_groupnorm_parts(n, h * w, c)
        ws = torch.empty((n, parts + 1, c, 2), dtype=torch.float32, device="cuda")
        p.keep.append(ws)
        p.add("groupnorm", L.his_groupnorm_act, x.ptr, n, h * w, c, x.cs, groups, p.const(gamma).data_ptr(), p.const(beta).data_ptr(), 1e-5, act,
              1.0, 1 if res is not None else 0, res.ptr if res is not None else None, res.cs if res is not None else 0, ws.data_ptr(), out.ptr,
              out.cs, 1 if split else 0)
        outs.append(out)
    p.replay(); torch.cuda.synchronize()
    xin, rin = x.torch_nchw().cpu(), r.torch_nchw().cpu()
    y = F.group_norm(xin, groups, gamma, beta, 1e-5)
    for out, ref in zip(outs, (F.relu(y + rin), F.silu(y))):
        assert (out.torch_nchw().cpu() - ref).abs().max() <= (2e-5 if split else 2e-3) * max(1.0, float(ref.abs().max()))


@SPLIT
def test_glue_kernels_match_torch(split):
    p = _plan(split); lib = p.lib
    S = 1 if split else 0
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(9)
    n, h, w, c = 3, 16, 12, 256
    x = p.act(n, h, w, c); x.fill_nhwc(torch.randn(n, h, w, c, generator=g))
    xr = x.torch_nchw().cpu()
    # maxpool
    mp = p.act(n, h // 2, w // 2, c)
    L.check(lib.his_maxpool2(x.ptr, n, h, w, c, x.cs, mp.ptr, mp.cs, S, st))
    # spatial attention 7x7
    wsa = torch.randn(1, 2, 7, 7, generator=g) * 0.2
    sa = p.act(n, h, w, c); stats = p.f32(n, h, w, 2)
    L.check(lib.his_spatial_attention(x.ptr, n, h, w, c, x.cs, p.const(wsa.reshape(2, 7, 7)).data_ptr(), 7, stats.data_ptr(), sa.ptr, sa.cs, S, st))
    # nearest resize into a slice, bilinear fp32, NHWC->NCHW
    cat = p.act(n, 31, 25, c + 8)
    L.check(lib.his_resize_nearest(x.ptr, n, h, w, c, x.cs, 31, 25, cat.slice(8, c).ptr, cat.cs, S, st))
    t = torch.randn(n, 2, 10, 14, generator=g)
    bl = p.f32(n, 2, 23, 17)
    L.check(lib.his_resize_bilinear_f32(p.const(t).data_ptr(), n * 2, 10, 14, 23, 17, bl.data_ptr(), st))
    nchw = p.f32(n, c, h, w)
    L.check(lib.his_nhwc_half_to_nchw_float(x.ptr, n, h * w, c, x.cs, nchw.data_ptr(), S, st))
    torch.cuda.synchronize()
    assert torch.equal(mp.torch_nchw().cpu(), F.max_pool2d(xr, 2))
    s = torch.cat([xr.mean(1, keepdim=True), xr.max(1, keepdim=True)[0]], 1)
    ref = xr * torch.sigmoid(F.conv2d(s, wsa, padding=3))
    assert (sa.torch_nchw().cpu() - ref).abs().max() <= (2e-5 if split else 2e-3) * float(ref.abs().max())
    assert torch.equal(cat.slice(8, c).torch_nchw().cpu(), F.interpolate(xr, size=(31, 25), mode="nearest"))
    assert (bl.cpu() - F.interpolate(t, size=(23, 17), mode="bilinear", align_corners=False)).abs().max() < 1e-5
    assert torch.equal(nchw.cpu(), xr)


def test_head_tail_kernels_match_torch():
    p = _plan(); lib = p.lib
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(2)
    n, h, w = 3, 8, 6
    low = torch.randn(n, 2, h, w, generator=g)
    wt = torch.randn(2, 32, 2, 2, generator=g) * 0.5
    s, t = torch.rand(32, generator=g) + 0.5, torch.randn(32, generator=g) * 0.1
    w1, b1 = torch.randn(2, 32, generator=g) * 0.3, torch.randn(2, generator=g) * 0.1
    out = p.f32(n, 2, 2 * h, 2 * w)
    L.check(lib.his_upsample_bgfg(p.const(low).data_ptr(), n, h, w, p.const(wt).data_ptr(), p.const(s).data_ptr(), p.const(t).data_ptr(),
                                  p.const(w1).data_ptr(), p.const(b1).data_ptr(), 1, 1.0, out.data_ptr(), st))
    tn = torch.randn(n, 2, 2 * h, 2 * w, generator=g)
    logits = p.f32(n, 3, 2 * h, 2 * w)
    L.check(lib.his_head_combine(out.data_ptr(), p.const(tn).data_ptr(), n, 2 * h, 2 * w, logits.data_ptr(), st))
    torch.cuda.synchronize()
    y = F.relu(F.conv_transpose2d(low, wt, stride=2) * s.view(1, -1, 1, 1) + t.view(1, -1, 1, 1))
    ref = F.conv2d(y, w1.view(2, 32, 1, 1), b1)
    assert (out.cpu() - ref).abs().max() < 1e-4
    fg = F.softmax(ref, 1)[:, 1]
    want = torch.stack([ref[:, 0], ref[:, 1] + tn[:, 0] * fg, ref[:, 1] + tn[:, 1] * fg], 1)
    assert (logits.cpu() - want).abs().max() < 1e-4


@SPLIT
def test_fused_roi_align_matches_reference_golden_and_plain_kernel(split):
    """his_roi_align_fused (both aligners in one launch, warp per (ROI, output row), source rows staged in shared memory) against
    the reference-generated goldens of DynamicRoIAlign (hed/dynamic_roi_align.py:56-171) and, on a wide image whose ROI rows exceed
    the staging capacity (direct-gather fallback), against the plain kernel -- identical results required."""
    import ctypes
    p = _plan(split); lib = p.lib
    st = torch.cuda.current_stream().cuda_stream
    g = common.golden("roi_align")
    feat, rois = g["feat"].cuda(), g["rois"].cuda()
    f0, f1 = feat[:, :2].contiguous(), feat[:, 2:].contiguous()
    n = rois.shape[0]
    for tag, (sh, sw), aligned, (oh, ow) in [("a640", (640.0, 640.0), 1, (16, 12)), ("ahw", (37.0, 53.0), 1, (16, 12)), ("u_hw", (37.0, 53.0), 0, (7, 9)),
                                            ("a64", (64.0, 64.0), 1, (5, 3))]:
        o0, o1 = p.f32(n, 2, oh, ow), p.f32(n, 3, oh, ow)
        h0, h1 = p.act(n, oh, ow, 16).slice(8, 2), p.act_zeroed(n, oh, ow, 3)
        L.check(lib.his_roi_align_fused(f0.data_ptr(), 2, sh, sw, aligned, h0.ptr, h0.cs, o0.data_ptr(),
                                        f1.data_ptr(), 3, sh, sw, aligned, h1.ptr, h1.cs, o1.data_ptr(),
                                        3, 37, 53, rois.data_ptr(), n, oh, ow, 1 if split else 0, st))
        torch.cuda.synchronize()
        got = torch.cat([o0, o1], 1).cpu()
        assert (got - g[tag]).abs().max() < 2e-5, tag
        tol = 1e-6 if split else 1e-3          # the NHWC slices hold the same values, rounded to fp16 (or to a hi + lo pair)
        assert (h0.torch_nchw().cpu() - o0.cpu()).abs().max() <= tol * max(1.0, float(g[tag].abs().max()))
        assert (h1.torch_nchw().cpu() - o1.cpu()).abs().max() <= tol * max(1.0, float(g[tag].abs().max()))
    # wide image: rows of up to 1000 source pixels (> staging capacity) next to narrow ones, one source only
    gen = torch.Generator().manual_seed(4)
    wide = torch.randn(2, 3, 20, 1000, generator=gen).cuda()
    r = torch.tensor([[0, 0.0, 0.1, 1.0, 0.9], [1, 0.2, 0.0, 0.45, 1.0], [1, 0.9, 0.5, 0.1, 0.6], [5, 0.1, 0.1, 0.5, 0.5]], dtype=torch.float32).cuda()
    a, b = p.f32(4, 3, 6, 40), p.f32(4, 3, 6, 40)
    L.check(lib.his_roi_align_fused(wide.data_ptr(), 3, 20.0, 1000.0, 1, None, 0, a.data_ptr(), None, 0, 0.0, 0.0, 0, None, 0, None,
                                    2, 20, 1000, r.data_ptr(), 4, 6, 40, 0, st))
    L.check(lib.his_roi_align(wide.data_ptr(), 0, 3 * 20 * 1000, 20 * 1000, 1000, 1, 2, 3, 20, 1000, r.data_ptr(), 4, 6, 40, 20.0, 1000.0, 1, None, 0,
                              b.data_ptr(), 0, st))
    torch.cuda.synchronize()
    assert torch.equal(a, b)

