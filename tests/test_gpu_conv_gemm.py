"""GPU: the tcgen05 implicit-GEMM convolution against a plain torch fp32 reference of the same op on the same
fp16-rounded operands (tolerance: fp32 accumulation-order noise + one fp16 output rounding)."""
import pytest
import torch
import torch.nn.functional as F

from human_instance_segmentation_b200 import engine
from human_instance_segmentation_b200.engine import ACT, RES_ADD, RES_MUL, RES_NONE

pytestmark = pytest.mark.gpu


def _act(y, code, beta):
    return {0: lambda v: v, 1: F.relu, 2: F.silu, 3: torch.sigmoid, 4: lambda v: v * torch.sigmoid(beta * v), 5: F.gelu}[code](y)


def run_case(n, h, w, cin, cout, k, act=1, res_mode=RES_NONE, transposed=False, in_slice=None, out_slice=None, beta=1.0, seed=0, split=False):
    """split=True: the split-fp16 ("strict") kernels -- operands hi + lo, reference on the UNROUNDED fp32 operands."""
    import ctypes
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(seed)
    plan = engine.Plan(dev, split)
    # input as a slice of a wider buffer when requested
    in_total, in_off = in_slice or (cin, 0)
    xbuf = plan.act(n, h, w, in_total)
    xbuf.buf.copy_((torch.randn(n, h, w, xbuf.cs, generator=g)).half())       # neighbours of the slice (and, split, its lo planes)
    x = xbuf.slice(in_off, cin)
    if split:
        x.fill_nhwc(torch.randn(n, h, w, cin, generator=g))
    if transposed:
        wt = torch.randn(cin, cout, 2, 2, generator=g) * (1.0 / cin) ** 0.5
    else:
        wt = torch.randn(cout, cin, k, k, generator=g) * (1.0 / (cin * k * k)) ** 0.5
    scale = torch.rand(cout, generator=g) + 0.5
    shift = torch.randn(cout, generator=g) * 0.1
    nt, bn = ctypes.c_int(), ctypes.c_int()
    plan.lib.his_conv_gemm_tile_n(cout, ctypes.byref(nt), ctypes.byref(bn))
    slab = nt.value * bn.value
    wp, cin_pad = engine.pack_gemm_weight(wt, slab, transposed, scale, split=split)   # BatchNorm scale folded into the weights
    oh, ow = (2 * h, 2 * w) if transposed else (h, w)
    out_total, out_off = out_slice or (cout, 0)
    obuf = plan.act(n, oh, ow, out_total)
    obuf.buf.fill_(7.0)
    out = obuf.slice(out_off, cout)
    res = None
    if res_mode != RES_NONE:
        res = plan.act(n, oh, ow, cout)
        res.fill_nhwc(torch.randn(n, oh, ow, cout, generator=g))
    plan.conv_gemm(x, plan.const(wp, torch.float16), cin_pad, plan.const(engine.pad_vec(shift, slab)),
                   out, k, act, beta, res, res_mode, transposed)
    plan.replay()
    torch.cuda.synchronize()
    got = out.torch_nchw().cpu()
    # reference
    xin = x.torch_nchw().cpu()
    # the kernel's operand: weights with the scale folded in, rounded to fp16
    wh = wt * (scale.view(1, -1, 1, 1) if transposed else scale.view(-1, 1, 1, 1))
    if not split:
        wh = wh.half().float()
    else:       # what the hi + lo pair can represent
        wh = wh.half().float() + (wh - wh.half().float()).half().float()
    if transposed:
        y = F.conv_transpose2d(xin, wh, stride=2)
    else:
        y = F.conv2d(xin, wh, padding=k // 2)
    y = y + shift.view(1, -1, 1, 1)
    if res_mode == RES_ADD:
        y = y + res.torch_nchw().cpu()
    y = _act(y, act, beta)
    if res_mode == RES_MUL:
        y = y * res.torch_nchw().cpu()
    err = (got - y).abs().max().item()
    ref = y.abs().max().item()
    # untouched channels of a sliced output buffer must keep their fill value
    if out_slice:
        full = obuf.buf.float().cpu()
        mask = torch.ones(full.shape[-1], dtype=torch.bool)
        mask[out_off:out_off + cout] = False
        if split:
            mask[obuf.cs // 2 + out_off:obuf.cs // 2 + out_off + cout] = False
        assert (full[..., mask] == 7.0).all(), "store spilled outside the output channel slice"
    return err, ref


CASES = [
    dict(n=2, h=64, w=48, cin=256, cout=256, k=3),                           # the hot shape
    dict(n=3, h=64, w=48, cin=64, cout=64, k=3, res_mode=RES_ADD),           # residual block tail
    dict(n=2, h=16, w=12, cin=128, cout=256, k=3),                           # small spatial (partial tiles)
    dict(n=2, h=20, w=15, cin=144, cout=288, k=3),                           # bc72: odd width, 2 N-tiles, K tail
    dict(n=1, h=80, w=60, cin=72, cout=72, k=3, res_mode=RES_ADD),           # channel counts not multiples of 16/64
    dict(n=2, h=64, w=48, cin=258, cout=256, k=1, act=0, in_slice=(258, 0)),  # feature_combiner: K tail of 2
    dict(n=2, h=64, w=48, cin=128, cout=256, k=1, act=3, res_mode=RES_MUL),  # fg_gate: sigmoid * shared
    dict(n=2, h=32, w=24, cin=256, cout=128, k=1, act=1, transposed=True),   # ConvT k2s2
    dict(n=2, h=16, w=12, cin=256, cout=128, k=1, act=0, transposed=True, out_slice=(256, 0)),   # ConvT into concat slice
    dict(n=2, h=32, w=24, cin=128, cout=64, k=3, act=1, out_slice=(128, 64)),  # skip written into upper concat slice
    dict(n=1, h=96, w=128, cin=32, cout=16, k=3),                            # smp decoder tail (narrow)
    dict(n=1, h=30, w=40, cin=432, cout=256, k=3),                           # smp decoder block 0 (K = 7 blocks, tail 48)
    dict(n=2, h=24, w=32, cin=16, cout=96, k=1, act=2),                      # MBConv expand, SiLU
    dict(n=1, h=12, w=16, cin=384, cout=768, k=3),                           # B7 bottleneck width: 3 N-tiles
    dict(n=5, h=64, w=48, cin=256, cout=64, k=3, act=4, beta=1.5),           # swish(beta)
    dict(n=1, h=7, w=5, cin=64, cout=32, k=3, act=5),                        # tiny image, gelu
    dict(n=2, h=48, w=64, cin=16, cout=16, k=3),                             # BK=16 path (SWIZZLE_32B operands)
    dict(n=2, h=30, w=40, cin=24, cout=144, k=1, act=2),                     # BK=32 path (SWIZZLE_64B), K tail
    dict(n=2, h=15, w=20, cin=40, cout=240, k=1, act=2),                     # BK=16, 3 K blocks
    dict(n=1, h=60, w=80, cin=96, cout=32, k=3, out_slice=(64, 32)),         # BK=32, 3 blocks per tap
    dict(n=2, h=30, w=40, cin=144, cout=24, k=1, act=0, res_mode=RES_ADD),   # MBConv project + skip, N=32 tile with 24 valid
    # halo mode (3x3): window loaded once per K block by cp.async, nine taps = nine descriptor offsets
    dict(n=3, h=80, w=60, cin=72, cout=72, k=3, res_mode=RES_ADD, out_slice=(144, 72)),   # clipped tiles + channel tail -> direct epilogue
    dict(n=2, h=33, w=21, cin=24, cout=40, k=3, act=2),                      # BK=32 with one half-empty K=16 step; ragged image
    dict(n=2, h=16, w=8, cin=320, cout=16, k=3, in_slice=(384, 64)),         # 5 K blocks, nine taps per weight stage, input slice
    dict(n=1, h=128, w=96, cin=128, cout=128, k=3, res_mode=RES_MUL, act=3),  # mask-resolution layer, gate operand
    dict(n=1, h=17, w=9, cin=8, cout=304, k=3),                               # Cin = 8 (one real plane), 2 N tiles
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items() if k in ("h", "w", "cin", "cout", "k")))
def test_conv_gemm_matches_torch_fp32(case):
    err, ref = run_case(**case)
    assert err <= 2e-3 * max(ref, 1.0), (err, ref)   # fp16 output rounding: 2^-11 relative


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items() if k in ("h", "w", "cin", "cout", "k")))
def test_conv_gemm_split_fp16_matches_torch_fp32(case):
    """The strict-precision kernels (three MMA passes over hi / lo operand planes): against torch fp32 on the operands the pairs
    represent -- the error left is the fp32 accumulation order, the approximate sigmoid (2 ulp) and the lo plane's own rounding."""
    err, ref = run_case(**case, split=True)
    assert err <= 2e-5 * max(ref, 1.0), (err, ref)


@pytest.mark.parametrize("case", [dict(n=3, h=64, w=48, cin=128, cout=128, k=3, res_mode=RES_ADD),
                                  dict(n=2, h=30, w=40, cin=200, cout=256, k=3, act=2),
                                  dict(n=1, h=16, w=12, cin=64, cout=160, k=3)],
                         ids=["n128-res", "n256-silu-clipped", "n160-small"])
def test_cta_pair_halo_variant_matches_torch_fp32(case, monkeypatch):
    """HIS_GEMM_PAIR_HALO: the cta_group::2 kernels fed by the halo window (off by default, kept as a tuning variant); the per-tap
    pair kernels are what CASES with Cout >= 128 and an even tile count run through by default."""
    monkeypatch.setenv("HIS_GEMM_PAIR_HALO", "128")
    err, ref = run_case(**case)
    assert err <= 2e-3 * max(ref, 1.0), (err, ref)


@pytest.mark.parametrize("split", [False, True], ids=["fp16", "split"])
def test_fused_nearest_upsample_concat_matches_torch(split):
    """his_conv_gemm_set_upsampled_input: channels [0, low_c) gathered from the half-resolution tensor at (y>>1, x>>1), the rest
    from the concat buffer == F.interpolate(nearest) + cat + conv3x3 (smp UnetDecoderBlock)."""
    import ctypes
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(7)
    for (n, h, w, low_c, skip_c, cout) in [(2, 24, 32, 64, 24, 64), (1, 34, 18, 32, 0, 16), (3, 16, 48, 128, 40, 96)]:
        plan = engine.Plan(dev, split)
        low = plan.act(n, h // 2, w // 2, low_c)
        low.fill_nhwc(torch.randn(n, h // 2, w // 2, low_c, generator=g))
        cat = plan.act(n, h, w, low_c + skip_c)
        cat.fill_nhwc(torch.randn(n, h, w, low_c + skip_c, generator=g))        # the first low_c channels are never read
        cin = low_c + skip_c
        wt = torch.randn(cout, cin, 3, 3, generator=g) * (1.0 / (cin * 9)) ** 0.5
        shift = torch.randn(cout, generator=g) * 0.1
        assert plan.lib.his_conv_gemm_can_fuse_upsample(h, w, cin, cout, low_c) == 1
        nt, bn = ctypes.c_int(), ctypes.c_int()
        plan.lib.his_conv_gemm_tile_n(cout, ctypes.byref(nt), ctypes.byref(bn))
        slab = nt.value * bn.value
        wp, cin_pad = engine.pack_gemm_weight(wt, slab, False, split=split)
        out = plan.act(n, h, w, cout)
        plan.conv_gemm(cat, plan.const(wp, torch.float16), cin_pad, plan.const(engine.pad_vec(shift, slab)), out, 3, ACT["relu"], up_input=low)
        plan.replay()
        torch.cuda.synchronize()
        x = torch.cat([F.interpolate(low.torch_nchw().cpu(), scale_factor=2, mode="nearest"), cat.torch_nchw().cpu()[:, low_c:]], 1)
        wref = wt.half().float() + ((wt - wt.half().float()).half().float() if split else 0.0)
        want = F.relu(F.conv2d(x, wref, padding=1) + shift.view(1, -1, 1, 1))
        err = (out.torch_nchw().cpu() - want).abs().max().item()
        assert err <= (2e-5 if split else 2e-3) * max(want.abs().max().item(), 1.0), (n, h, w, low_c, skip_c, cout, err)


def test_row_scale_and_channel_statistics_epilogue():
    """his_conv_gemm_set_row_ops: y = act(row_scale[pix]*conv(x) + shift (+res)); stats[pix] = (mean_c y, max_c y) -- the two halves
    of the SpatialAttentionModule fusion (attention_modules.py:67-113)."""
    import ctypes
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(11)
    n, h, w, cin, cout = 3, 16, 12, 256, 256
    plan = engine.Plan(dev)
    x = plan.act(n, h, w, cin); x.buf.copy_(torch.randn(x.buf.shape, generator=g).half())
    res = plan.act(n, h, w, cout); res.buf.copy_(torch.randn(res.buf.shape, generator=g).half())
    wt = torch.randn(cout, cin, 3, 3, generator=g) * (1.0 / (cin * 9)) ** 0.5
    shift = torch.randn(cout, generator=g) * 0.1
    nt, bn = ctypes.c_int(), ctypes.c_int()
    plan.lib.his_conv_gemm_tile_n(cout, ctypes.byref(nt), ctypes.byref(bn))
    wp, cin_pad = engine.pack_gemm_weight(wt, nt.value * bn.value, False)
    out = plan.act(n, h, w, cout)
    stats = plan.f32(n, h, w, 2)
    plan.conv_gemm(x, plan.const(wp, torch.float16), cin_pad, plan.const(engine.pad_vec(shift, cout)), out, 3, ACT["relu"], 1.0, res, RES_ADD, stats_out=stats)
    # consumer: ConvT k2s2 of the gated tensor == row scale inside the transposed conv
    gate = torch.rand(n, h, w, generator=g)
    wT = torch.randn(cout, 128, 2, 2, generator=g) * (1.0 / cout) ** 0.5
    shiftT = torch.randn(128, generator=g) * 0.1
    wpT, cin_padT = engine.pack_gemm_weight(wT, 128, True)
    outT = plan.act(n, 2 * h, 2 * w, 128)
    plan.conv_gemm(out, plan.const(wpT, torch.float16), cin_padT, plan.const(shiftT), outT, 1, ACT["relu"], transposed=True, row_scale=plan.const(gate))
    plan.replay()
    torch.cuda.synchronize()
    y = F.relu(F.conv2d(x.torch_nchw().cpu(), wt.half().float(), padding=1) + shift.view(1, -1, 1, 1) + res.torch_nchw().cpu())
    assert (stats.cpu()[..., 0] - y.mean(1)).abs().max() < 2e-3 and (stats.cpu()[..., 1] - y.max(1)[0]).abs().max() < 5e-3
    yh = out.torch_nchw().cpu()
    wantT = F.relu(F.conv_transpose2d(yh * gate[:, None], wT.half().float(), stride=2) + shiftT.view(1, -1, 1, 1))
    errT = (outT.torch_nchw().cpu() - wantT).abs().max().item()
    assert errT <= 3e-3 * max(wantT.abs().max().item(), 1.0), errT
