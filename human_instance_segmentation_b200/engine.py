"""Launch-plan builder/executor for the B200 path.

A ``Plan`` is a flat list of ``(c_function, args)`` records bound to pre-allocated HBM buffers
(activations NHWC fp16, tails NCHW fp32).  It is built once per input geometry
``(B, H, W, n_rois)`` from the model's parameters (BatchNorm folded, weights repacked to the
kernels' layouts) and replayed every forward on the caller's current CUDA stream -- optionally
captured into a CUDA graph.  torch is used for memory and streams only.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Tuple

import torch

from . import lib as _lib

ACT = {"none": 0, "relu": 1, "silu": 2, "sigmoid": 3, "swish": 4, "gelu": 5}
RES_NONE, RES_ADD, RES_MUL = 0, 1, 2


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class Act:
    """A channel slice [c_off, c_off+C) of an NHWC fp16 buffer [N,H,W,cs].

    ``split`` (the "strict" precision mode): every value is a pair of fp16 numbers x = hi + lo stored [hi channels | lo channels]
    inside one pixel -- the buffer is [N,H,W,cs] with the hi planes in [0, cs/2) and the lo plane of a channel cs/2 elements
    after its hi plane; ``cs`` stays the pixel stride the kernels are given."""

    __slots__ = ("buf", "N", "H", "W", "C", "cs", "c_off", "zero_tail", "split", "ln_ws")

    def __init__(self, buf: torch.Tensor, C: int, c_off: int = 0, split: bool = False):
        assert buf.dtype == torch.float16 and buf.dim() == 4 and buf.is_contiguous()
        self.buf = buf
        self.N, self.H, self.W, self.cs = buf.shape
        self.C, self.c_off = C, c_off
        self.split = bool(split)
        self.zero_tail = False          # True: channels [C, cs) are zero and nobody ever writes them (see Plan.act_zeroed)
        self.ln_ws = None               # (double [N, parts, 2], parts): LayerNorm2d partial statistics written by the producing GEMM
        assert c_off % 8 == 0 and self.cs % (16 if split else 8) == 0 and c_off + C <= self.width

    @property
    def width(self) -> int:
        """Channel capacity of the buffer (per plane)."""
        return self.cs // 2 if self.split else self.cs

    @property
    def ptr(self) -> int:
        return self.buf.data_ptr() + 2 * self.c_off

    def slice(self, c_off: int, C: int) -> "Act":
        return Act(self.buf, C, self.c_off + c_off, self.split)

    def widen(self, C: int) -> "Act":
        """The first C channels of the underlying buffer (a concat buffer read as one tensor)."""
        return Act(self.buf, C, 0, self.split)

    def fill_nhwc(self, t: torch.Tensor):
        """Debug / test helper: writes fp32 values [N,H,W,C] into the slice (hi = fp16(t), lo = fp16(t - hi) when split)."""
        t = t.to(self.buf.device, torch.float32)
        hi = t.half()
        self.buf[..., self.c_off:self.c_off + self.C] = hi
        if self.split:
            lo = self.cs // 2
            self.buf[..., lo + self.c_off:lo + self.c_off + self.C] = (t - hi.float()).half()

    def torch_nchw(self) -> torch.Tensor:
        """Debug view (fp32 NCHW copy through torch) -- tests only."""
        v = self.buf[..., self.c_off:self.c_off + self.C].float()
        if self.split:
            lo = self.cs // 2
            v = v + self.buf[..., lo + self.c_off:lo + self.c_off + self.C].float()
        return v.permute(0, 3, 1, 2).contiguous()


class NullAct(Act):
    """Geometry-only stand-in for an activation that is never written (GEMM with a fused tail and store_main=0)."""

    def __init__(self, buf: torch.Tensor, N: int, H: int, W: int, C: int):
        self.buf, self.N, self.H, self.W, self.C = buf, N, H, W, C
        self.cs, self.c_off = round_up(C, 16), 0
        self.zero_tail = False
        self.split = False
        self.ln_ws = None


class Plan:
    def __init__(self, device: torch.device, split: bool = False):
        self.device = device
        self.split = bool(split)            # split-fp16 ("strict" precision) activations and weights
        self.lib = _lib.load()
        self.ops: List[Tuple] = []          # (name, cfunc, args-without-stream)
        self.keep: List[object] = []        # tensors / gemm plans the ops point into
        self.gemm_plans: List[ctypes.c_void_p] = []
        self.gemm_shapes: List[Tuple] = []
        self.flops = 0                      # algorithmic FLOPs (2*MAC, true channel counts) of one replay
        self.flops_by_tag: Dict[str, int] = {}
        self.tensor_flops = 0               # the part issued on tcgen05
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches = 0
        self.op_flops: List[int] = []
        self.op_desc: List[str] = []
        self.tag = ""

    # -------------------------------------------------------------- allocation
    def act(self, N, H, W, C, cs=None) -> Act:
        cs = cs or round_up(C, 8)
        buf = torch.empty((N, H, W, cs * (2 if self.split else 1)), dtype=torch.float16, device=self.device)
        self.keep.append(buf)
        return Act(buf, C, 0, self.split)

    def act_zeroed(self, N, H, W, C) -> Act:
        """Activation whose channel tail [C, cs) is zero for the plan's lifetime: producers write exactly C channels."""
        buf = torch.zeros((N, H, W, round_up(C, 8) * (2 if self.split else 1)), dtype=torch.float16, device=self.device)
        self.keep.append(buf)
        a = Act(buf, C, 0, self.split)
        a.zero_tail = True
        return a

    def null_act(self, N, H, W, C) -> NullAct:
        buf = torch.zeros(64, dtype=torch.float16, device=self.device)
        self.keep.append(buf)
        return NullAct(buf, N, H, W, C)

    def f32(self, *shape, zero=False) -> torch.Tensor:
        t = (torch.zeros if zero else torch.empty)(shape, dtype=torch.float32, device=self.device)
        self.keep.append(t)
        return t

    def const(self, t: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
        t = t.detach().to(device=self.device, dtype=dtype).contiguous()
        self.keep.append(t)
        return t

    def _add_flops(self, f: int, tensor: bool):
        self.flops += f
        self.flops_by_tag[self.tag] = self.flops_by_tag.get(self.tag, 0) + f
        if tensor:
            self.tensor_flops += f

    # kernels launched per op (cudaMemsetAsync is not one of ours)
    KERNELS_PER_OP = {"spatial_attention": 2, "input_affine": 3, "memset": 0, "se_gate": 3, "boundary_edges": 3, "layernorm2d": 2,
                      "groupnorm": 3}

    def add(self, name: str, fn, *args, flops: int = 0, desc: str = ""):
        self.ops.append((name, fn, args))
        self.op_flops.append(flops)
        self.op_desc.append(f"{self.tag}:{name} {desc}")
        self.launches += self.KERNELS_PER_OP.get(name, 1)

    # -------------------------------------------------------------- ops
    def conv_gemm(self, x: Act, w_packed: torch.Tensor, cin_pad: int, shift: torch.Tensor, out: Act,
                  ksize: int, act: int, beta: float = 1.0, res: Optional[Act] = None, res_mode: int = RES_NONE,
                  transposed: bool = False, tail=None, aux_f32: Optional[torch.Tensor] = None, in_gate: Optional[torch.Tensor] = None,
                  row_scale: Optional[torch.Tensor] = None, stats_out: Optional[torch.Tensor] = None, up_input: Optional[Act] = None,
                  res_scale: Optional[torch.Tensor] = None, alg_flops: Optional[int] = None, ln_stats: bool = False):
        """tail = (tail_w fp32 [tc, cout_slab], (b0, b1), tc, sigmoid?, out_f32 NCHW, store_main) fuses a 1x1 conv to <=2
        channels into the epilogue (his_conv_gemm_set_tail).  aux_f32: fp32 NCHW copy of the output written from the
        epilogue (his_conv_gemm_set_aux).  in_gate = (gate fp32 [N, Cin], fp16 scratch >= N*rows*cin_pad): per-image
        input-channel gate, folded into per-image weights by a small kernel ahead of the GEMM (his_scale_weights +
        his_conv_gemm_set_image_weights); the scratch may be shared by ops that run back to back on the stream."""
        L = self.lib
        h = ctypes.c_void_p()
        if transposed:
            assert out.H == 2 * x.H and out.W == 2 * x.W
        else:
            assert (out.N, out.H, out.W) == (x.N, x.H, x.W)
        _lib.check(L.his_conv_gemm_create(ctypes.byref(h), x.ptr, x.N, x.H, x.W, x.C, x.cs, w_packed.data_ptr(), cin_pad,
                                          out.ptr, out.C, out.cs, res.ptr if res is not None else None,
                                          res.cs if res is not None else 0, shift.data_ptr(), ksize,
                                          1 if transposed else 0, act, beta, res_mode, 1 if self.split else 0), "his_conv_gemm_create")
        self.gemm_plans.append(h)
        self.keep += [w_packed, shift]
        taps = 4 if transposed else ksize * ksize
        # algorithmic FLOPs of the REFERENCE convolution this launch computes (alg_flops: a re-formulated layer whose weight
        # matrix holds structural zeros reports the original conv's count, not the padded one)
        f = alg_flops if alg_flops is not None else 2 * x.N * x.H * x.W * x.C * out.C * taps
        if tail is not None:
            tw, (b0, b1), tc, sig, tout, store_main = tail
            _lib.check(L.his_conv_gemm_set_tail(h, tw.data_ptr(), float(b0), float(b1), tc, 1 if sig else 0, tout.data_ptr(),
                                                1 if store_main else 0), "his_conv_gemm_set_tail")
            self.keep += [tw, tout]
            f += 2 * x.N * x.H * x.W * out.C * tc
        if aux_f32 is not None:
            _lib.check(L.his_conv_gemm_set_aux(h, aux_f32.data_ptr()), "his_conv_gemm_set_aux")
            self.keep.append(aux_f32)
        if row_scale is not None or stats_out is not None:
            _lib.check(L.his_conv_gemm_set_row_ops(h, row_scale.data_ptr() if row_scale is not None else None,
                                                   stats_out.data_ptr() if stats_out is not None else None), "his_conv_gemm_set_row_ops")
            self.keep += [t for t in (row_scale, stats_out) if t is not None]
        if res_scale is not None:
            _lib.check(L.his_conv_gemm_set_res_scale(h, res_scale.data_ptr()), "his_conv_gemm_set_res_scale")
            self.keep.append(res_scale)
        if ln_stats:                      # LayerNorm2d statistics of the output from the epilogue (one (sum, sumsq) pair per work item)
            n_work = L.his_conv_gemm_work_items(h)
            ws = torch.zeros((max(n_work, 1), 2), dtype=torch.float64, device=self.device)
            parts = ctypes.c_int()
            _lib.check(L.his_conv_gemm_set_ln_partials(h, ws.data_ptr(), ctypes.byref(parts)), "his_conv_gemm_set_ln_partials")
            self.keep.append(ws)
            out.ln_ws = (ws, parts.value)
        if up_input is not None:          # channels [0, up_input.C) gathered from the half-resolution tensor (nearest 2x)
            assert (2 * up_input.H, 2 * up_input.W) == (x.H, x.W) and up_input.N == x.N
            _lib.check(L.his_conv_gemm_set_upsampled_input(h, up_input.ptr, up_input.C, up_input.cs), "his_conv_gemm_set_upsampled_input")
            self.keep.append(up_input.buf)
        if in_gate is not None:
            gate, wimg = in_gate
            rows = w_packed.numel() // cin_pad
            assert wimg.dtype == torch.float16 and wimg.numel() >= x.N * rows * cin_pad
            _lib.check(L.his_conv_gemm_set_image_weights(h, wimg.data_ptr()), "his_conv_gemm_set_image_weights")
            self.add("scale_weights", L.his_scale_weights, w_packed.data_ptr(), gate.data_ptr(), x.N, rows, cin_pad, x.C, wimg.data_ptr(),
                     1 if self.split else 0)
            self.keep += [gate, wimg]
        self._add_flops(f, True)
        self.add("conv_gemm", L.his_conv_gemm_run, h, flops=f,
                 desc=f"N{x.N} {x.H}x{x.W} cin{x.C} cout{out.C} k{ksize}{' T' if transposed else ''}{' res%d' % res_mode if res_mode else ''}"
                      f"{' +tail%d' % tail[2] if tail is not None else ''}")
        self.gemm_shapes.append((x.N, x.H, x.W, x.C, out.C, ksize, int(transposed)))

    def conv_direct(self, x, in_fmt: int, N, H, W, cin, in_cs, w: torch.Tensor, scale, shift, cout, k, stride, pad, act, beta=1.0,
                    in_affine: Optional[torch.Tensor] = None, res: Optional[Act] = None, res_mode: int = RES_NONE,
                    out: Optional[Act] = None, out_f32: Optional[torch.Tensor] = None):
        L = self.lib
        xptr = x.ptr if isinstance(x, Act) else x.data_ptr()
        self.keep += [w, scale, shift]
        ho, wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        fd = 2 * N * ho * wo * cin * cout * k * k
        self._add_flops(fd, False)
        self.add("conv_direct", L.his_conv_direct, xptr, in_fmt, in_affine.data_ptr() if in_affine is not None else None, N, H, W,
                 cin, in_cs, w.data_ptr(), scale.data_ptr(), shift.data_ptr(), cout, k, k, stride, pad, act, beta, res_mode,
                 res.ptr if res is not None else None, res.cs if res is not None else 0,
                 out.ptr if out is not None else None, out.cs if out is not None else 0,
                 out_f32.data_ptr() if out_f32 is not None else None, 1 if self.split else 0, flops=fd,
                 desc=f"N{N} {H}x{W} cin{cin} cout{cout} k{k} s{stride} fmt{in_fmt}")

    # -------------------------------------------------------------- execution
    def run(self, stream_ptr: int):
        s = ctypes.c_void_p(stream_ptr)
        for name, fn, args in self.ops:
            rc = fn(*args, s)
            if rc != 0:
                _lib.check(rc, name)

    def run_timed(self) -> List[Tuple[str, float, int]]:
        """Instrumented replay: every op bracketed by CUDA events on the launching stream.
        Returns [(op name, milliseconds, algorithmic flops)]; used by bench.py for the roofline figures."""
        stream = torch.cuda.current_stream(self.device)
        s = ctypes.c_void_p(stream.cuda_stream)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(self.ops) + 1)]
        evs[0].record(stream)
        for i, (name, fn, args) in enumerate(self.ops):
            rc = fn(*args, s)
            if rc != 0:
                _lib.check(rc, name)
            evs[i + 1].record(stream)
        stream.synchronize()
        return [(self.ops[i][0], evs[i].elapsed_time(evs[i + 1]), self.op_flops[i]) for i in range(len(self.ops))]

    def replay(self):
        """Runs the plan on the current torch stream (through a CUDA graph once captured)."""
        if self.graph is not None:
            self.graph.replay()
        else:
            self.run(torch.cuda.current_stream(self.device).cuda_stream)

    def capture(self):
        """Captures the launch list into a CUDA graph (launch-bound tails become one submission)."""
        torch.cuda.synchronize(self.device)
        side = torch.cuda.Stream(self.device)
        with torch.cuda.stream(side):
            self.run(side.cuda_stream)          # warm-up outside capture (module load, attribute sets)
        side.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            self.run(torch.cuda.current_stream(self.device).cuda_stream)
        self.graph = g

    def __del__(self):
        try:
            for h in self.gemm_plans:
                self.lib.his_conv_gemm_destroy(h)
        except Exception:
            pass


# ----------------------------------------------------------------------------- weight packing
def fold_bn(conv_bias: Optional[torch.Tensor], bn, cout: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """y = conv(x)*scale + shift  ==  BN_eval(conv(x) + bias)   (BatchNorm2d eps from the module)."""
    bias = conv_bias.detach().float().cpu() if conv_bias is not None else torch.zeros(cout)
    bn = getattr(bn, "batch_norm", bn)          # MixedNormalization in eval mode is its BatchNorm2d
    if bn is None:
        return torch.ones(cout), bias
    g, b = bn.weight.detach().float().cpu(), bn.bias.detach().float().cpu()
    m, v = bn.running_mean.detach().float().cpu(), bn.running_var.detach().float().cpu()
    scale = g / torch.sqrt(v + bn.eps)
    return scale, b - m * scale + bias * scale


def pad_vec(v: torch.Tensor, n: int) -> torch.Tensor:
    out = torch.zeros(n, dtype=torch.float32)
    out[: v.numel()] = v
    return out


def pack_gemm_weight(w: torch.Tensor, cout_slab: int, transposed: bool = False, scale: Optional[torch.Tensor] = None,
                     split: bool = False) -> Tuple[torch.Tensor, int]:
    """Conv2d weight [Cout,Cin,kh,kw] -> fp16 [1][taps][cout_slab][cin_pad];
    ConvTranspose2d(k2,s2) weight [Cin,Cout,2,2] -> fp16 [4 groups (dy,dx)][1][cout_slab][cin_pad].
    ``scale`` [Cout] (the folded BatchNorm scale) multiplies the fp32 weights per output channel before the fp16 rounding.
    ``split``: rows become [W_hi | W_lo] (W_hi = fp16(w), W_lo = fp16(w - W_hi)); the returned cin_pad counts both halves."""
    if split:
        hi, k1 = pack_gemm_weight(w, cout_slab, transposed, scale)
        full = _packed_f32(w, cout_slab, transposed, scale, k1)
        lo = (full - hi.float()).half()
        return torch.cat([hi, lo], dim=-1).contiguous(), 2 * k1
    w = w.detach().float().cpu()
    if scale is not None:
        sc = scale.detach().float().cpu()
        w = w * (sc.view(1, -1, 1, 1) if transposed else sc.view(-1, 1, 1, 1))
        if float(w.abs().max()) > 6.0e4:
            raise _lib.HisError("folded conv*BatchNorm weight exceeds the fp16 range (|w*gamma/sqrt(var+eps)| > 6e4)")
    # K is zero-padded to a multiple of 64 so that every weight TMA box lies inside the tensor (a clipped box takes a far
    # slower path through the TMA unit)
    if transposed:
        cin, cout = w.shape[0], w.shape[1]
        cin_pad = round_up(cin, 64)
        out = torch.zeros(4, 1, cout_slab, cin_pad, dtype=torch.float16)
        for dy in range(2):
            for dx in range(2):
                out[dy * 2 + dx, 0, :cout, :cin] = w[:, :, dy, dx].t().half()
        return out.contiguous(), cin_pad
    cout, cin, kh, kw = w.shape
    cin_pad = round_up(cin, 64)
    out = torch.zeros(1, kh * kw, cout_slab, cin_pad, dtype=torch.float16)
    out[0, :, :cout, :cin] = w.permute(2, 3, 0, 1).reshape(kh * kw, cout, cin).half()
    return out.contiguous(), cin_pad


def _packed_f32(w: torch.Tensor, cout_slab: int, transposed: bool, scale: Optional[torch.Tensor], cin_pad: int) -> torch.Tensor:
    """The packed layout of ``pack_gemm_weight`` in fp32 (before any rounding)."""
    w = w.detach().float().cpu()
    if scale is not None:
        sc = scale.detach().float().cpu()
        w = w * (sc.view(1, -1, 1, 1) if transposed else sc.view(-1, 1, 1, 1))
    if transposed:
        cin, cout = w.shape[0], w.shape[1]
        out = torch.zeros(4, 1, cout_slab, cin_pad)
        for dy in range(2):
            for dx in range(2):
                out[dy * 2 + dx, 0, :cout, :cin] = w[:, :, dy, dx].t()
        return out
    cout, cin, kh, kw = w.shape
    out = torch.zeros(1, kh * kw, cout_slab, cin_pad)
    out[0, :, :cout, :cin] = w.permute(2, 3, 0, 1).reshape(kh * kw, cout, cin)
    return out


def pack_direct_weight(w: torch.Tensor, f32: bool = False) -> torch.Tensor:
    """[Cout,Cin,kh,kw] -> [kh][kw][Cin][Cout], fp16 (fp32 for the split-fp16 "strict" mode)."""
    t = w.detach().float().cpu().permute(2, 3, 1, 0).contiguous()
    return t if f32 else t.half()
