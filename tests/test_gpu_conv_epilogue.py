"""GPU: the epilogue variants of the tcgen05 conv kernel (fused 1x1 tail, fp32 NCHW export, residual add / gate, several N tiles)
at sizes where every persistent CTA drains many tiles -- the regime of bench.py, where the hand-over of the staging buffers
between the epilogue warps, the TMA stores and the residual loads is exercised back to back (a tail-only layer whose warps
were allowed to drift apart hung at that size while every small case passed).  Reference: torch fp32 on the same fp16 operands.
"""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from human_instance_segmentation_b200 import engine
from human_instance_segmentation_b200.engine import ACT, RES_ADD, RES_MUL, RES_NONE

pytestmark = pytest.mark.gpu


def _ref_act(y, code):
    return {0: lambda v: v, 1: F.relu, 2: F.silu, 3: torch.sigmoid}[code](y)


def run(n, h, w, cin, cout, k, act=1, res_mode=RES_NONE, tail_c=0, tail_sigmoid=False, aux=False, rsc=False, seed=0):
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(seed)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        plan = engine.Plan(dev)
        x = plan.act(n, h, w, cin)
        x.buf.copy_(torch.randn(x.buf.shape, generator=g).half())
        wt = torch.randn(cout, cin, k, k, generator=g) * (1.0 / (cin * k * k)) ** 0.5
        shift = torch.randn(cout, generator=g) * 0.1
        nt, bn = ctypes.c_int(), ctypes.c_int()
        plan.lib.his_conv_gemm_tile_n(cout, ctypes.byref(nt), ctypes.byref(bn))
        slab = nt.value * bn.value
        wp, cin_pad = engine.pack_gemm_weight(wt, slab, False)
        res = None
        if res_mode != RES_NONE:
            res = plan.act(n, h, w, cout)
            res.buf.copy_(torch.randn(res.buf.shape, generator=g).half())
        tail = tout = None
        if tail_c:
            tw = torch.zeros(tail_c, slab)
            tw[:, :cout] = torch.randn(tail_c, cout, generator=g) * (1.0 / cout) ** 0.5
            tb = (torch.randn(2, generator=g) * 0.1).tolist()
            tout = plan.f32(n, tail_c, h, w)
            tail = (plan.const(tw), (tb[0], tb[1]), tail_c, tail_sigmoid, tout, False)          # tail-only: the main output is not stored
            out = plan.null_act(n, h, w, cout)
        else:
            out = plan.act(n, h, w, cout)
        auxt = plan.f32(n, cout, h, w) if aux else None
        rscale = plan.const(torch.rand(n, cout, generator=g) + 0.5) if rsc else None       # per-(image, channel) scale of the residual operand
        plan.conv_gemm(x, plan.const(wp, torch.float16), cin_pad, plan.const(engine.pad_vec(shift, slab)), out, k, act, 1.0, res, res_mode,
                       tail=tail, aux_f32=auxt, res_scale=rscale)
        for _ in range(3):          # replays: the buffers' parities continue across launches of the same plan
            plan.replay()
        torch.cuda.synchronize()
        y = F.conv2d(x.torch_nchw(), wt.half().float().to(dev), padding=k // 2) + shift.to(dev).view(1, -1, 1, 1)
        r = res.torch_nchw() * (rscale.view(n, cout, 1, 1) if rsc else 1.0) if res is not None else None
        if res_mode == RES_ADD:
            y = y + r
        y = _ref_act(y, act)
        gate = y
        if res_mode == RES_MUL:
            y = y * r
        scale = max(y.abs().max().item(), 1.0)
        if tail_c:
            t = torch.einsum("nchw,oc->nohw", y, tw[:, :cout].to(dev)) + torch.tensor(tb[:tail_c], device=dev).view(1, -1, 1, 1)
            if tail_sigmoid:
                t = torch.sigmoid(t)
            err = (tout - t).abs().max().item()
            assert err <= 2e-4 * max(t.abs().max().item(), 1.0), ("tail", err)
        else:
            err = (out.torch_nchw() - y).abs().max().item()
            assert err <= 2e-3 * scale, ("main", err, scale)
        if aux:         # the export holds the fp32 value before the fp16 rounding (the gate itself for RES_MUL)
            err = (auxt - gate).abs().max().item()
            assert err <= 2e-4 * max(gate.abs().max().item(), 1.0), ("aux", err)
    finally:
        torch.backends.cudnn.allow_tf32 = old


CASES = {
    "tail2-res-n128-pair": dict(n=48, h=64, w=48, cin=128, cout=128, k=3, res_mode=RES_ADD, tail_c=2),
    "tail1-n64-halo": dict(n=64, h=64, w=48, cin=64, cout=64, k=3, tail_c=1, tail_sigmoid=True),
    "tail2-res-n32-halo": dict(n=64, h=64, w=48, cin=64, cout=32, k=3, res_mode=RES_ADD, tail_c=2),
    "tail2-res-rsc-n128-pair": dict(n=48, h=64, w=48, cin=128, cout=128, k=3, res_mode=RES_ADD, tail_c=2, rsc=True),
    "res-rsc-k3-n256": dict(n=24, h=64, w=48, cin=256, cout=256, k=3, res_mode=RES_ADD, rsc=True),
    "tail1-k1-n256": dict(n=64, h=32, w=24, cin=64, cout=256, k=1, tail_c=1),            # split 1x1 layer goes back to one N tile
    "aux-gate-k1": dict(n=64, h=64, w=48, cin=128, cout=256, k=1, act=3, res_mode=RES_MUL, aux=True),
    "aux-res-k3-pair": dict(n=24, h=64, w=48, cin=256, cout=256, k=3, res_mode=RES_ADD, aux=True),
    "aux-clipped": dict(n=9, h=60, w=44, cin=64, cout=72, k=3, aux=True),
    "res-k3-n256": dict(n=32, h=64, w=48, cin=128, cout=256, k=3, res_mode=RES_ADD),
    "res-k1-n96": dict(n=64, h=60, w=80, cin=240, cout=40, k=1, act=0, res_mode=RES_ADD),
    "ntiles3-silu-k1": dict(n=64, h=30, w=40, cin=112, cout=672, k=1, act=2),
    "plain-k1-n256": dict(n=64, h=64, w=48, cin=256, cout=256, k=1),
    "plain-ragged": dict(n=33, h=50, w=37, cin=48, cout=80, k=3, res_mode=RES_ADD),
}


@pytest.mark.parametrize("name", list(CASES))
def test_epilogue_variant_many_tiles_per_cta(name):
    run(**CASES[name])


@pytest.mark.parametrize("flags", ["32", "64", "96", "128"])
def test_general_chunk_body_and_top_of_chunk_sync_still_match(flags, monkeypatch):
    """HIS_GEMM_DEBUG 32 / 64: the general (branchy) chunk body and the wait + barrier at the top of every chunk -- the forms the
    straight-line body and the lean hand-over replaced; kept selectable for A/B timing, so kept correct."""
    monkeypatch.setenv("HIS_GEMM_DEBUG", flags)
    run(**CASES["tail2-res-n128-pair"])
    run(**CASES["aux-gate-k1"])
    run(**CASES["res-k3-n256"])
    run(**CASES["tail2-res-rsc-n128-pair"])
