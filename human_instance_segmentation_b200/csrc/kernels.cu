// CUDA-core kernels of the path: Dynamic RoI Align gather, direct convolutions for the odd
// shapes (Cin = 2/3, Cout <= 2 tails, stem, segmentation head), depthwise conv + BN + SiLU,
// squeeze-excite / channel attention, spatial attention, pooling, resizes and the hierarchical
// combine.  All are HBM-bound: coalesced channel-last accesses, 16-byte vectors where the channel
// count allows, one pass per tensor.  Dense convolutions live in conv_gemm_sm100.cu.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

thread_local char g_err[256] = "";

constexpr int kThreads = 256;

inline int grid_for(long long work, int per_block = kThreads) {
  long long g = (work + per_block - 1) / per_block;
  if (g > 148LL * 64) g = 148LL * 64;   // grid-stride loops; multiple of the SM count
  if (g < 1) g = 1;
  return (int)g;
}

// Reductions over pixels keep one 8-channel group per thread: with 256-thread blocks that holds when the grid
// stride gridDim.x*256 is a multiple of the number of groups, i.e. gridDim.x is a multiple of cgs/gcd(cgs,256).
inline int gcd_int(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }
inline int threads_multiple_of(int cgs) { return cgs > 0 ? kThreads : 0; }
inline int grid_multiple(int cgs, int threads = kThreads) { return cgs / gcd_int(cgs, threads); }

// ------------------------------------------------------------------------------------ RoI Align
// reference hed/dynamic_roi_align.py:56-171 (see oracle/headport.py:roi_align for the derivation).
__device__ __forceinline__ float linspace01(int i, int n) {
  if (n <= 1) return 0.0f;
  const float step = __fdiv_rn(1.0f, (float)(n - 1));
  return (i < n / 2) ? __fmul_rn(step, (float)i) : __fsub_rn(1.0f, __fmul_rn(step, (float)(n - 1 - i)));
}

__device__ __forceinline__ float roi_px(float lo, float hi, float g, int size, int aligned) {
  const float f = __fadd_rn(lo, __fmul_rn(g, __fsub_rn(hi, lo)));
  if (aligned) {
    const float nrm = __fsub_rn(__fmul_rn(__fdiv_rn(f, (float)(size - 1)), 2.0f), 1.0f);
    return __fmul_rn(__fdiv_rn(__fadd_rn(nrm, 1.0f), 2.0f), (float)(size - 1));
  }
  const float nrm = __fsub_rn(__fmul_rn(__fdiv_rn(f, (float)size), 2.0f), 1.0f);
  return __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(nrm, 1.0f), (float)size), 1.0f), 2.0f);
}

template <typename T>
__device__ __forceinline__ float ld_as_float(const T* p);
template <> __device__ __forceinline__ float ld_as_float<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_as_float<__half>(const __half* p) { return __half2float(__ldg(p)); }

// One thread per (roi, oy, ox); channels looped (C is 2 or 3 on the hot path).  Consecutive threads
// walk ox, so the four taps of a warp fall on at most a few image rows (coalesced when the ROI is
// sampled near unit stride).
template <typename T>
__global__ void roi_align_kernel(const T* __restrict__ feat, long long sN, long long sC, long long sH, long long sW, int B, int C,
                                 int H, int W, const float* __restrict__ rois, int n_rois, int oh, int ow, float scale_h,
                                 float scale_w, int aligned, __half* __restrict__ out_h, int out_cs, int out_lo, float* __restrict__ out_f) {
  const long long total = (long long)n_rois * oh * ow;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % ow);
    const int oy = (int)((idx / ow) % oh);
    const int k = (int)(idx / ((long long)ow * oh));
    const float* r = rois + 5 * k;
    const int b = (int)r[0];
    const float x1 = __fmul_rn(r[1], scale_w), y1 = __fmul_rn(r[2], scale_h);
    const float x2 = __fmul_rn(r[3], scale_w), y2 = __fmul_rn(r[4], scale_h);
    const float px = roi_px(x1, x2, linspace01(ox, ow), W, aligned);
    const float py = roi_px(y1, y2, linspace01(oy, oh), H, aligned);
    const float fx0 = floorf(px), fy0 = floorf(py);
    const int ix = (int)fx0, iy = (int)fy0;
    const float wx1 = px - fx0, wx0 = (fx0 + 1.0f) - px;
    const float wy1 = py - fy0, wy0 = (fy0 + 1.0f) - py;
    const bool okb = (b >= 0 && b < B) && isfinite(px) && isfinite(py);
    const bool x0ok = okb && ix >= 0 && ix < W, x1ok = okb && ix + 1 >= 0 && ix + 1 < W;
    const bool y0ok = iy >= 0 && iy < H, y1ok = iy + 1 >= 0 && iy + 1 < H;
    for (int c = 0; c < C; ++c) {
      const T* base = feat + (long long)(okb ? b : 0) * sN + (long long)c * sC;
      float v = 0.0f;
      if (x0ok && y0ok) v += ld_as_float(base + iy * sH + ix * sW) * (wx0 * wy0);
      if (x1ok && y0ok) v += ld_as_float(base + iy * sH + (ix + 1) * sW) * (wx1 * wy0);
      if (x0ok && y1ok) v += ld_as_float(base + (iy + 1) * sH + ix * sW) * (wx0 * wy1);
      if (x1ok && y1ok) v += ld_as_float(base + (iy + 1) * sH + (ix + 1) * sW) * (wx1 * wy1);
      if (out_h) his_st1(out_h + idx * out_cs + c, out_lo, v);
      if (out_f) out_f[(((long long)k * C + c) * oh + oy) * ow + ox] = v;
    }
  }
}

// Both aligners of the model in ONE launch (rgb.py:751-755: roi_align_mask on the 2-channel UNet logits, roi_align_rgb on the
// image), one WARP per (ROI, output row): an output row samples two source rows (iy, iy+1) over the ROI's x extent, so the warp
// stages exactly that segment of every channel in shared memory with coalesced 128-byte row reads (each source pixel is read
// once, instead of up to two 4-byte sector-sized gathers per output tap) and then every lane interpolates its output pixels
// from the staged rows, in the same operation order as roi_align_kernel (identical results).  Outputs go straight into the
// consumers' layouts: the NHWC fp16 slices (mask logits -> channels 256..257 of the feature_combiner input, RGB patches -> the
// 3->64 conv's input) and the NCHW fp32 aux tensors.  Segments longer than kRoiSeg pixels fall back to direct gathers.
constexpr int kRoiSeg = 296, kRoiWarps = 4, kRoiMaxC = 5;
struct RoiSrc {
  const float* feat; int C;            // [B,C,H,W] fp32 contiguous
  float scale_h, scale_w; int aligned;
  __half* out_h; int out_cs, out_lo;   // NHWC fp16 slice (may be null)
  float* out_f;                        // NCHW fp32 [n_rois,C,oh,ow] (may be null)
};
__global__ void __launch_bounds__(kRoiWarps * 32) roi_align_rows_kernel(RoiSrc s0, RoiSrc s1, int B, int H, int W, const float* __restrict__ rois,
                                                                        int n_rois, int oh, int ow) {
  __shared__ float s_seg[kRoiWarps][3][2][kRoiSeg];      // one source at a time (<= 3 channels): 7 KB per warp, 8 CTAs per SM
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row_id = (long long)blockIdx.x * kRoiWarps + warp;
  if (row_id >= (long long)n_rois * oh) return;
  const int k = (int)(row_id / oh), oy = (int)(row_id - (long long)k * oh);
  const float* r = rois + 5 * k;
  const int b = (int)r[0];
#pragma unroll
  for (int si = 0; si < 2; ++si) {
    const RoiSrc& s = si == 0 ? s0 : s1;
    if (s.C == 0) continue;
    const float x1 = __fmul_rn(r[1], s.scale_w), y1 = __fmul_rn(r[2], s.scale_h);
    const float x2 = __fmul_rn(r[3], s.scale_w), y2 = __fmul_rn(r[4], s.scale_h);
    const float py = roi_px(y1, y2, linspace01(oy, oh), H, s.aligned);
    const float fy0 = floorf(py);
    const int iy = (int)fy0;
    const float wy1 = py - fy0, wy0 = (fy0 + 1.0f) - py;
    const bool okb = (b >= 0 && b < B) && isfinite(py);
    const bool y0ok = okb && iy >= 0 && iy < H, y1ok = okb && iy + 1 >= 0 && iy + 1 < H;
    // x extent of the row's samples (the grid is monotonic in ox; either direction)
    const float pxa = roi_px(x1, x2, linspace01(0, ow), W, s.aligned), pxb = roi_px(x1, x2, linspace01(ow - 1, ow), W, s.aligned);
    const bool finite_x = isfinite(pxa) && isfinite(pxb);
    int lo = 0, hi = -1;
    if (finite_x) {
      lo = max((int)floorf(fminf(pxa, pxb)), 0);
      hi = min((int)floorf(fmaxf(pxa, pxb)) + 1, W - 1);
    }
    const bool staged = finite_x && (hi - lo + 1) <= kRoiSeg;
    const float* img = s.feat + (long long)(okb ? b : 0) * s.C * H * W;
    if (staged) {
      // all 2*C row segments of one 32-pixel column block are loaded before any is stored: 2*C independent loads in flight per
      // lane instead of one load -> store round trip per element
      const float* src0 = img + (long long)iy * W + lo;
      for (int i = lane; i <= hi - lo; i += 32) {
        float t[3][2];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int rr = 0; rr < 2; ++rr)
            t[c][rr] = (c < s.C && (rr == 0 ? y0ok : y1ok)) ? __ldg(src0 + ((long long)c * H + rr) * W + i) : 0.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c)
          if (c < s.C) { s_seg[warp][c][0][i] = t[c][0]; s_seg[warp][c][1][i] = t[c][1]; }
      }
    }
    __syncwarp();
    for (int ox = lane; ox < ow; ox += 32) {
      const float px = roi_px(x1, x2, linspace01(ox, ow), W, s.aligned);
      const float fx0 = floorf(px);
      const int ix = (int)fx0;
      const float wx1 = px - fx0, wx0 = (fx0 + 1.0f) - px;
      const bool okx = isfinite(px);
      const bool x0ok = okx && ix >= 0 && ix < W, x1ok = okx && ix + 1 >= 0 && ix + 1 < W;
      const long long opix = ((long long)k * oh + oy) * ow + ox;
      for (int c = 0; c < s.C; ++c) {
        float v = 0.0f;
        if (staged) {
          const float* t0 = s_seg[warp][c][0];
          const float* t1 = s_seg[warp][c][1];
          if (x0ok && y0ok) v += t0[ix - lo] * (wx0 * wy0);
          if (x1ok && y0ok) v += t0[ix + 1 - lo] * (wx1 * wy0);
          if (x0ok && y1ok) v += t1[ix - lo] * (wx0 * wy1);
          if (x1ok && y1ok) v += t1[ix + 1 - lo] * (wx1 * wy1);
        } else {
          const float* base = img + (long long)c * H * W;
          if (x0ok && y0ok) v += __ldg(base + (long long)iy * W + ix) * (wx0 * wy0);
          if (x1ok && y0ok) v += __ldg(base + (long long)iy * W + ix + 1) * (wx1 * wy0);
          if (x0ok && y1ok) v += __ldg(base + (long long)(iy + 1) * W + ix) * (wx0 * wy1);
          if (x1ok && y1ok) v += __ldg(base + (long long)(iy + 1) * W + ix + 1) * (wx1 * wy1);
        }
        if (s.out_h) his_st1(s.out_h + opix * s.out_cs + c, s.out_lo, v);
        if (s.out_f) s.out_f[(((long long)k * s.C + c) * oh + oy) * ow + ox] = v;
      }
    }
    __syncwarp();          // the staging rows are reused by the next source
  }
}

// ------------------------------------------------------------------------------------ direct conv
struct DirectConvParams {
  const void* in; int in_fmt;          // 0: NHWC fp16 (stride in_cs), 1: NCHW fp32
  const float* in_affine;              // optional [2*Cin]: x*a[c]+b[c] on in-bounds samples (input normalisation)
  int N, H, W, Cin, in_cs;
  const void* w;                       // [kh][kw][Cin][Cout], fp16 (fp32 when w_f32: the split-fp16 "strict" mode)
  const float* scale; const float* shift;
  int Cout, kh, kw, stride, pad, Ho, Wo;
  int act; float act_beta; int res_mode;
  const __half* res; int res_cs;
  __half* out_h; int out_cs;           // NHWC fp16 slice (may be null)
  float* out_f;                        // NCHW fp32 (may be null)
  int w_f32, in_lo, res_lo, out_lo;    // split-fp16 activations: offset of the lo plane (0: plain fp16)
};
__device__ __forceinline__ float dc_w(const DirectConvParams& p, long long i) {
  return p.w_f32 ? __ldg((const float*)p.w + i) : __half2float(__ldg((const __half*)p.w + i));
}

template <int COT>
__global__ void direct_conv_kernel(const DirectConvParams p) {
  const int groups = p.Cout / COT;
  const long long total = (long long)p.N * p.Ho * p.Wo * groups;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(idx % groups);
    long long pix = idx / groups;
    const int ox = (int)(pix % p.Wo), oy = (int)((pix / p.Wo) % p.Ho), n = (int)(pix / ((long long)p.Wo * p.Ho));
    const int co = cg * COT;
    float acc[COT];
#pragma unroll
    for (int t = 0; t < COT; ++t) acc[t] = 0.0f;
    for (int ky = 0; ky < p.kh; ++ky) {
      const int iy = oy * p.stride - p.pad + ky;
      if (iy < 0 || iy >= p.H) continue;
      for (int kx = 0; kx < p.kw; ++kx) {
        const int ix = ox * p.stride - p.pad + kx;
        if (ix < 0 || ix >= p.W) continue;
        const long long wrow = ((long long)(ky * p.kw + kx) * p.Cin) * p.Cout + co;
        if (p.in_fmt == 0) {
          const __half* ip = (const __half*)p.in + ((long long)(n * p.H + iy) * p.W + ix) * p.in_cs;
          for (int ci = 0; ci < p.Cin; ++ci) {
            const float x = his_ld1(ip + ci, p.in_lo);
            const long long wp = wrow + (long long)ci * p.Cout;
#pragma unroll
            for (int t = 0; t < COT; ++t) acc[t] = fmaf(x, dc_w(p, wp + t), acc[t]);
          }
        } else {
          const float* ip = (const float*)p.in + (long long)n * p.Cin * p.H * p.W + (long long)iy * p.W + ix;
          for (int ci = 0; ci < p.Cin; ++ci) {
            float x = __ldg(ip + (long long)ci * p.H * p.W);
            if (p.in_affine) x = fmaf(x, __ldg(p.in_affine + ci), __ldg(p.in_affine + p.Cin + ci));
            const long long wp = wrow + (long long)ci * p.Cout;
#pragma unroll
            for (int t = 0; t < COT; ++t) acc[t] = fmaf(x, dc_w(p, wp + t), acc[t]);
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < COT; ++t) {
      float y = acc[t] * __ldg(p.scale + co + t) + __ldg(p.shift + co + t);
      float r = 0.0f;
      if (p.res_mode) r = his_ld1(p.res + pix * p.res_cs + co + t, p.res_lo);
      if (p.res_mode == HIS_RES_ADD) y += r;
      y = his_act(y, p.act, p.act_beta);
      if (p.res_mode == HIS_RES_MUL) y *= r;
      acc[t] = y;
    }
    if (p.out_h) {
      __half* op = p.out_h + pix * p.out_cs + co;
#pragma unroll
      for (int t = 0; t < COT; ++t) his_st1(op + t, p.out_lo, acc[t]);
    }
    if (p.out_f) {
#pragma unroll
      for (int t = 0; t < COT; ++t) p.out_f[(((long long)n * p.Cout + co + t) * p.Ho + oy) * p.Wo + ox] = acc[t];
    }
  }
}

// Small-Cin convolution (Cin <= 4: UNet stem 3->32 s2 from NCHW fp32, ROI feature extractor 3->64, fg_gate 2->64):
// one thread = one output pixel x 32 output channels; the k*k*Cin input window sits in registers, weights are staged
// in shared memory as fp32 and read as broadcast float4s.
template <int KK /*k*k*/, int CIN>
__global__ void small_cin_conv_kernel(const DirectConvParams p) {
  extern __shared__ float s_w[];            // [KK*CIN][Cout] weights, then scale[Cout], shift[Cout] (broadcast reads instead of
                                            // two global loads per output channel per thread)
  const int nw = KK * CIN * p.Cout;
  float* s_sc = s_w + nw;
  float* s_sh = s_sc + p.Cout;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) s_w[i] = dc_w(p, i);
  for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) { s_sc[i] = __ldg(p.scale + i); s_sh[i] = __ldg(p.shift + i); }
  __syncthreads();
  const int groups = p.Cout / 32;
  const long long total = (long long)p.N * p.Ho * p.Wo * groups;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(idx % groups);
    const long long pix = idx / groups;
    const int ox = (int)(pix % p.Wo), oy = (int)((pix / p.Wo) % p.Ho), n = (int)(pix / ((long long)p.Wo * p.Ho));
    float xin[KK * CIN];
#pragma unroll
    for (int t = 0; t < KK; ++t) {
      const int ky = t / p.kw, kx = t - ky * p.kw;
      const int iy = oy * p.stride - p.pad + ky, ix = ox * p.stride - p.pad + kx;
      const bool ok = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        float v = 0.0f;
        if (ok) {
          if (p.in_fmt == 0) v = his_ld1((const __half*)p.in + ((long long)(n * p.H + iy) * p.W + ix) * p.in_cs + ci, p.in_lo);
          else {
            v = __ldg((const float*)p.in + ((long long)(n * CIN + ci) * p.H + iy) * p.W + ix);
            if (p.in_affine) v = fmaf(v, __ldg(p.in_affine + ci), __ldg(p.in_affine + CIN + ci));
          }
        }
        xin[t * CIN + ci] = v;
      }
    }
    const int co = cg * 32;
    float acc[32];
#pragma unroll
    for (int t = 0; t < 32; ++t) acc[t] = 0.0f;
#pragma unroll
    for (int j = 0; j < KK * CIN; ++j) {
      const float4* wr = reinterpret_cast<const float4*>(s_w + j * p.Cout + co);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 wv = wr[q];
        acc[4 * q] = fmaf(xin[j], wv.x, acc[4 * q]); acc[4 * q + 1] = fmaf(xin[j], wv.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(xin[j], wv.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(xin[j], wv.w, acc[4 * q + 3]);
      }
    }
    __half* op = p.out_h + pix * p.out_cs + co;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float y8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = q * 8 + e;
        y8[e] = his_act(fmaf(acc[c], s_sc[co + c], s_sh[co + c]), p.act, p.act_beta);
      }
      his_st8(op + q * 8, p.out_lo, y8);
    }
  }
}

// Small-Cout convolution from an NHWC fp16 slice (smp segmentation head 3x3 16->1, un-fused 1x1 tails): one thread =
// one output pixel, 16-byte channel vectors, fp32 weights in shared memory, fp32 NCHW output.
template <int COUT>
__global__ void small_cout_conv_kernel(const DirectConvParams p) {
  extern __shared__ float s_w[];            // [kh*kw][Cin][COUT]
  const int nw = p.kh * p.kw * p.Cin * COUT;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) s_w[i] = dc_w(p, i);
  __syncthreads();
  const long long total = (long long)p.N * p.Ho * p.Wo;
  for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < total; pix += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(pix % p.Wo), oy = (int)((pix / p.Wo) % p.Ho), n = (int)(pix / ((long long)p.Wo * p.Ho));
    float acc[COUT];
#pragma unroll
    for (int t = 0; t < COUT; ++t) acc[t] = 0.0f;
    for (int ky = 0; ky < p.kh; ++ky) {
      const int iy = oy * p.stride - p.pad + ky;
      if (iy < 0 || iy >= p.H) continue;
      for (int kx = 0; kx < p.kw; ++kx) {
        const int ix = ox * p.stride - p.pad + kx;
        if (ix < 0 || ix >= p.W) continue;
        const __half* ip = (const __half*)p.in + ((long long)(n * p.H + iy) * p.W + ix) * p.in_cs;
        const float* wr = s_w + (long long)(ky * p.kw + kx) * p.Cin * COUT;
        for (int c8 = 0; c8 < p.Cin; c8 += 8) {
          float xf[8];
          his_ld8(ip + c8, p.in_lo, xf);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
#pragma unroll
            for (int t = 0; t < COUT; ++t) acc[t] = fmaf(xf[e], wr[(c8 + e) * COUT + t], acc[t]);
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < COUT; ++t) {
      const float y = his_act(acc[t] * __ldg(p.scale + t) + __ldg(p.shift + t), p.act, p.act_beta);
      p.out_f[(((long long)n * COUT + t) * p.Ho + oy) * p.Wo + ox] = y;
    }
  }
}

// 3x3 stride-1 pad-1 conv to ONE output channel (the smp segmentation head 16->1 at full image resolution).  A warp walks down a
// 30-column strip: lane L owns input column x0-1+L and loads every input pixel exactly once (32 lanes x 32 B = one contiguous
// 1 KB row segment); per loaded row it adds that pixel's dot products with the nine taps into the three output rows the row
// touches, and a finished output row is assembled from the neighbours' partial sums with two shuffles:
//   out[y][x] = sum_ky ( x[y+ky-1][x-1].w[ky][0] + x[y+ky-1][x].w[ky][1] + x[y+ky-1][x+1].w[ky][2] ).
// fp32 weights as broadcast reads from shared memory, fp32 NCHW output (coalesced over the lanes).
constexpr int kHeadStripCols = 30, kHeadStripRows = 32;
template <int CIN>
__global__ void __launch_bounds__(kThreads) head3x3_c1_kernel(const DirectConvParams p) {
  __shared__ __align__(16) float s_w[9 * CIN];
  for (int i = threadIdx.x; i < 9 * CIN; i += blockDim.x) s_w[i] = dc_w(p, i);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int ncol = (p.W + kHeadStripCols - 1) / kHeadStripCols, nstrip = (p.H + kHeadStripRows - 1) / kHeadStripRows;
  const long long total = (long long)p.N * nstrip * ncol;
  const float sc = __ldg(p.scale), sh = __ldg(p.shift);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long item = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5); item < total; item += nwarps) {
    const int cgp = (int)(item % ncol), strip = (int)((item / ncol) % nstrip), n = (int)(item / ((long long)ncol * nstrip));
    const int x = cgp * kHeadStripCols - 1 + lane;
    const bool xin_ok = x >= 0 && x < p.W;
    const int y0 = strip * kHeadStripRows, y1 = min(y0 + kHeadStripRows, p.H);
    float P[3], Q[3] = {0.0f, 0.0f, 0.0f}, R[3] = {0.0f, 0.0f, 0.0f};    // partial sums of output rows r+1, r, r-1
    for (int r = y0 - 1; r <= y1; ++r) {
      float xin[CIN];
      if (xin_ok && r >= 0 && r < p.H) {
        // in_fmt 2: phase-packed input [N, H/2, W/2, 4*CIN] -- pixel (r, x) is channel block (r&1)*2 + (x&1) of low pixel (r>>1, x>>1)
        const __half* px = p.in_fmt == 2
            ? (const __half*)p.in + ((long long)(n * (p.H >> 1) + (r >> 1)) * (p.W >> 1) + (x >> 1)) * p.in_cs + (((r & 1) << 1) | (x & 1)) * CIN
            : (const __half*)p.in + ((long long)(n * p.H + r) * p.W + x) * p.in_cs;
#pragma unroll
        for (int c8 = 0; c8 < CIN; c8 += 8) his_ld8(px + c8, p.in_lo, xin + c8);
      } else {
#pragma unroll
        for (int c = 0; c < CIN; ++c) xin[c] = 0.0f;
      }
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
#pragma unroll
        for (int c4 = 0; c4 < CIN; c4 += 4) {
          const float4 w0 = *reinterpret_cast<const float4*>(s_w + (0 * 3 + kx) * CIN + c4);
          const float4 w1 = *reinterpret_cast<const float4*>(s_w + (1 * 3 + kx) * CIN + c4);
          const float4 w2 = *reinterpret_cast<const float4*>(s_w + (2 * 3 + kx) * CIN + c4);
          a0 = fmaf(xin[c4], w0.x, a0); a0 = fmaf(xin[c4 + 1], w0.y, a0); a0 = fmaf(xin[c4 + 2], w0.z, a0); a0 = fmaf(xin[c4 + 3], w0.w, a0);
          a1 = fmaf(xin[c4], w1.x, a1); a1 = fmaf(xin[c4 + 1], w1.y, a1); a1 = fmaf(xin[c4 + 2], w1.z, a1); a1 = fmaf(xin[c4 + 3], w1.w, a1);
          a2 = fmaf(xin[c4], w2.x, a2); a2 = fmaf(xin[c4 + 1], w2.y, a2); a2 = fmaf(xin[c4 + 2], w2.z, a2); a2 = fmaf(xin[c4 + 3], w2.w, a2);
        }
        P[kx] = a0; Q[kx] += a1; R[kx] += a2;        // input row r is tap ky = 0 / 1 / 2 of output rows r+1 / r / r-1
      }
      // output row r-1 is complete: the kx = 0 term comes from the left neighbour's column, the kx = 2 term from the right one
      const float left = __shfl_up_sync(0xffffffffu, R[0], 1), right = __shfl_down_sync(0xffffffffu, R[2], 1);
      const int o = r - 1;
      if (o >= y0 && lane >= 1 && lane <= kHeadStripCols && xin_ok)
        p.out_f[((long long)n * p.H + o) * p.W + x] = his_act((left + R[1] + right) * sc + sh, p.act, p.act_beta);
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) { R[kx] = Q[kx]; Q[kx] = P[kx]; }
    }
  }
}

// ------------------------------------------------------------------------------------ depthwise conv
// NHWC fp16, 8 channels (16 B) per thread, fused BN scale/shift + activation, optional per-(n,c)
// sums for the squeeze-excite pooling that follows (fp32 atomics, one per thread per channel after
// a warp-level merge of threads that share the channel group is not possible in general -> the
// block first reduces into shared memory).
struct DwParams {
  const __half* in; int N, H, W, C, in_cs;
  const __half* w;            // [k*k][C]
  const float* scale; const float* shift;
  int k, stride, pad, Ho, Wo, act;
  __half* out; int out_cs;
  float* pool;                // [N][gridDim.x][C] per-block partial sums of the activated output (may be null)
};

template <int ACT>
__device__ __forceinline__ float act_ct(float v) {
  if (ACT == HIS_ACT_RELU) return fmaxf(v, 0.0f);
  if (ACT == HIS_ACT_SILU) return v * his_sigmoid_fast(v, HIS_NEG_LOG2E);
  if (ACT == HIS_ACT_SIGMOID) return his_sigmoid_fast(v, HIS_NEG_LOG2E);
  return v;
}

// One thread = one VEC-channel group x one output column x a vertical strip of RY output rows.  The K*K filter taps of
// the thread's channel group live in registers (the group is fixed per thread), and every input row of the strip is
// loaded once (K vectors) and applied to all output rows it overlaps, so a thread issues K*((RY-1)*S+K)/RY loads per
// output instead of 2*K*K (taps + weights): the kernel leaves the L1-bandwidth bound of the one-output-per-thread form.
// VEC = 8 (16-byte vectors) for 3x3, 4 (8-byte vectors) for 5x5 so that 25 taps fit the register file.
constexpr int kDwThreads = 128;
template <int VEC> struct DwVec;
template <> struct DwVec<8> { typedef uint4 T; };
template <> struct DwVec<4> { typedef uint2 T; };

template <int K, int S, int ACT, int VEC, int RY>
__global__ void __launch_bounds__(kDwThreads, 3) depthwise_kernel(const DwParams p) {
  typedef typename DwVec<VEC>::T V;
  constexpr int RIN = (RY - 1) * S + K;
  extern __shared__ float s_pool[];     // [blockDim.x][VEC] per-thread sums, reduced in a fixed order (deterministic)
  const int cgs = p.C / VEC;
  const int n = blockIdx.y;
  const int strips = (p.Ho + RY - 1) / RY;
  const int per_img = strips * p.Wo * cgs;
  // fixed channel group per thread: the grid stride gridDim.x*blockDim.x is a multiple of cgs (host guarantees)
  const int cg = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) % cgs);
  const int c0 = cg * VEC;
  float psum[VEC], sc[VEC], sh[VEC];
  __half2 wreg[K * K][VEC / 2];
#pragma unroll
  for (int e = 0; e < VEC; ++e) { psum[e] = 0.0f; sc[e] = __ldg(p.scale + c0 + e); sh[e] = __ldg(p.shift + c0 + e); }
#pragma unroll
  for (int t = 0; t < K * K; ++t) {
    const V wv = __ldg(reinterpret_cast<const V*>(p.w + (long long)t * p.C + c0));
#pragma unroll
    for (int e = 0; e < VEC / 2; ++e) wreg[t][e] = reinterpret_cast<const __half2*>(&wv)[e];
  }
  const __half* inb = p.in + (long long)n * p.H * p.W * p.in_cs + c0;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < per_img; idx += gridDim.x * blockDim.x) {
    const int t2 = idx / cgs;
    const int strip = t2 / p.Wo, ox = t2 - strip * p.Wo;
    const int oy0 = strip * RY;
    const int iy0 = oy0 * S - p.pad, ix0 = ox * S - p.pad;
    float acc[RY][VEC];
#pragma unroll
    for (int j = 0; j < RY; ++j)
#pragma unroll
      for (int e = 0; e < VEC; ++e) acc[j][e] = 0.0f;
#pragma unroll
    for (int r = 0; r < RIN; ++r) {
      const int iy = iy0 + r;
      const bool rowok = iy >= 0 && iy < p.H;
      const int iyc = min(max(iy, 0), p.H - 1);
      V xv[K];
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {      // clamped (always valid) address; out-of-image taps are zeroed below
        const int ixc = min(max(ix0 + kx, 0), p.W - 1);
        xv[kx] = __ldg(reinterpret_cast<const V*>(inb + ((long long)iyc * p.W + ixc) * p.in_cs));
      }
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const int ix = ix0 + kx;
        const bool ok = rowok && ix >= 0 && ix < p.W;
        float xf[VEC];
#pragma unroll
        for (int e = 0; e < VEC / 2; ++e) {
          const float2 f = __half22float2(reinterpret_cast<const __half2*>(&xv[kx])[e]);
          xf[2 * e] = ok ? f.x : 0.0f; xf[2 * e + 1] = ok ? f.y : 0.0f;
        }
#pragma unroll
        for (int j = 0; j < RY; ++j) {
          const int ky = r - j * S;         // compile-time after unrolling
          if (ky >= 0 && ky < K) {
#pragma unroll
            for (int e = 0; e < VEC / 2; ++e) {
              const float2 wf = __half22float2(wreg[ky * K + kx][e]);
              acc[j][2 * e] = fmaf(xf[2 * e], wf.x, acc[j][2 * e]);
              acc[j][2 * e + 1] = fmaf(xf[2 * e + 1], wf.y, acc[j][2 * e + 1]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < RY; ++j) {
      const int oy = oy0 + j;
      if (oy < p.Ho) {
        V ov;
        __half2* o = reinterpret_cast<__half2*>(&ov);
#pragma unroll
        for (int e = 0; e < VEC; e += 2) {
          const float y0 = act_ct<ACT>(acc[j][e] * sc[e] + sh[e]);
          const float y1 = act_ct<ACT>(acc[j][e + 1] * sc[e + 1] + sh[e + 1]);
          o[e >> 1] = __floats2half2_rn(y0, y1);
          // pool what the next layer will actually read (the fp16-rounded value)
          const float2 rf = __half22float2(o[e >> 1]);
          psum[e] += rf.x; psum[e + 1] += rf.y;
        }
        *reinterpret_cast<V*>(p.out + ((long long)(n * p.Ho + oy) * p.Wo + ox) * p.out_cs + c0) = ov;
      }
    }
  }
  if (p.pool) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) s_pool[threadIdx.x * VEC + e] = psum[e];
    __syncthreads();
    float* dst = p.pool + ((long long)n * gridDim.x + blockIdx.x) * p.C;
    const int first = (int)((blockIdx.x * (long long)blockDim.x) % cgs);   // channel group of thread 0
    for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
      const int g2 = c / VEC, e = c % VEC;
      float s = 0.0f;
      for (int t = (g2 - first + cgs) % cgs; t < (int)blockDim.x; t += cgs) s += s_pool[t * VEC + e];
      dst[c] = s;
    }
  }
}

// ---- shared-memory tiled depthwise conv (the one the path runs).  CTA = 8 output rows x (8 strips of XPT columns) x 32 channels:
// the input window of the tile is staged once in shared memory (80-byte pixel pitch: the four 16-byte channel vectors of a pixel
// plus 16 B of padding, so the strips of a quarter-warp fall in distinct banks), the K*K x 32 filter taps sit next to it, and a
// thread produces XPT adjacent outputs of one 8-channel vector: a window row is read once (NX vectors) for all XPT outputs.
// 256 threads, <= 64 registers of state -> 2-3 CTAs per SM instead of the 12 warps of the register-resident form above.
// Per-(image, tile) channel sums of the fp16-rounded outputs go to pool[n][tile][c] in a fixed order (deterministic SE pooling).
constexpr int kDw2Threads = 256, kDw2Rows = 8, kDw2Strips = 8, kDw2Cb = 32, kDw2Pitch = 80;

template <int K, int S, int XPT>
struct Dw2Cfg {
  static constexpr int TOW = kDw2Strips * XPT, TOH = kDw2Rows;
  static constexpr int IW = (TOW - 1) * S + K, IH = (TOH - 1) * S + K;
  static constexpr int NX = (XPT - 1) * S + K;                       // window vectors a thread reads per filter row
  static constexpr int kInBytes = IW * IH * kDw2Pitch;
  static constexpr int kWBytes = K * K * kDw2Cb * 4;                  // filter taps as fp32 (converted once per CTA)
  static constexpr int kSmem = kInBytes + kWBytes + kDw2Threads * 8 * 4;   // + pooling scratch
};

template <int K, int S, int XPT, int ACT>
__global__ void __launch_bounds__(kDw2Threads) depthwise_tiled_kernel(const DwParams p, int tiles_x) {
  using Cfg = Dw2Cfg<K, S, XPT>;
  extern __shared__ __align__(16) unsigned char dw_smem[];
  unsigned char* s_in = dw_smem;
  float* s_w = reinterpret_cast<float*>(dw_smem + Cfg::kInBytes);
  float* s_pool = reinterpret_cast<float*>(dw_smem + Cfg::kInBytes + Cfg::kWBytes);
  const int n = blockIdx.z, cb0 = blockIdx.y * kDw2Cb;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const int oy0 = ty * Cfg::TOH, ox0 = tx * Cfg::TOW;
  const int iy0 = oy0 * S - p.pad, ix0 = ox0 * S - p.pad;
  const int nvec = min(4, (p.C - cb0) >> 3);                          // valid 8-channel vectors of this channel block
  // ---- stage the window (zero padding resolved here) and the filter taps
  const __half* inb = p.in + (long long)n * p.H * p.W * p.in_cs + cb0;
  // cp.async (zero-filling form): every load of the window is in flight at once, no register round trip
  const uint32_t s_in_addr = (uint32_t)__cvta_generic_to_shared(s_in);
  for (int i = threadIdx.x; i < Cfg::IW * Cfg::IH * 4; i += kDw2Threads) {
    const int v = i & 3, px = i >> 2;
    const int ly = px / Cfg::IW, lx = px - ly * Cfg::IW;
    const int iy = iy0 + ly, ix = ix0 + lx;
    const bool ok = v < nvec && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
    const __half* src = ok ? inb + ((long long)iy * p.W + ix) * p.in_cs + v * 8 : p.in;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s_in_addr + (uint32_t)(px * kDw2Pitch + v * 16)), "l"(src), "r"(ok ? 16u : 0u) : "memory");
  }
  for (int i = threadIdx.x; i < K * K * 4; i += kDw2Threads) {
    const int v = i & 3, t = i >> 2;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (v < nvec) val = __ldg(reinterpret_cast<const uint4*>(p.w + (long long)t * p.C + cb0 + v * 8));
    const __half2* wh = reinterpret_cast<const __half2*>(&val);
    float2* dst = reinterpret_cast<float2*>(s_w + t * kDw2Cb + v * 8);
#pragma unroll
    for (int e = 0; e < 4; ++e) dst[e] = __half22float2(wh[e]);
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  // ---- compute: lane -> (vector v, strip), warp -> output row
  const int v = threadIdx.x & 3, strip = (threadIdx.x >> 2) & 7, row = threadIdx.x >> 5;
  const int c0 = cb0 + v * 8;
  // accumulators, inputs and taps as fp32 PAIRS: one FFMA2 (fma.rn.f32x2, same rounding as two FFMAs) per two channels halves the
  // FMA issue slots of this issue-bound kernel
  float2 acc2[XPT][4];
#pragma unroll
  for (int j = 0; j < XPT; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc2[j][e] = make_float2(0.0f, 0.0f);
  const unsigned char* base = s_in + ((row * S) * Cfg::IW + strip * XPT * S) * kDw2Pitch + v * 16;
#pragma unroll
  for (int ky = 0; ky < K; ++ky) {
    float2 xf[Cfg::NX][4];
#pragma unroll
    for (int i = 0; i < Cfg::NX; ++i) {
      const uint4 xv = *reinterpret_cast<const uint4*>(base + (ky * Cfg::IW + i) * kDw2Pitch);
      const __half2* xh = reinterpret_cast<const __half2*>(&xv);
#pragma unroll
      for (int e = 0; e < 4; ++e) xf[i][e] = __half22float2(xh[e]);
    }
#pragma unroll
    for (int kx = 0; kx < K; ++kx) {
      const float4 w0 = *reinterpret_cast<const float4*>(s_w + (ky * K + kx) * kDw2Cb + v * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(s_w + (ky * K + kx) * kDw2Cb + v * 8 + 4);
      const float2 wf[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y), make_float2(w1.z, w1.w)};
#pragma unroll
      for (int j = 0; j < XPT; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc2[j][e] = his_ffma2(xf[j * S + kx][e], wf[e], acc2[j][e]);
    }
  }
  float acc[XPT][8];
#pragma unroll
  for (int j = 0; j < XPT; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) { acc[j][2 * e] = acc2[j][e].x; acc[j][2 * e + 1] = acc2[j][e].y; }
  // ---- BN + activation, store, per-thread channel sums of what the next layer will read
  float psum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int oy = oy0 + row;
  if (v < nvec && oy < p.Ho) {
    float sc[8], sh[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { sc[e] = __ldg(p.scale + c0 + e); sh[e] = __ldg(p.shift + c0 + e); }
#pragma unroll
    for (int j = 0; j < XPT; ++j) {
      const int ox = ox0 + strip * XPT + j;
      if (ox < p.Wo) {
        uint4 ov;
        __half2* o = reinterpret_cast<__half2*>(&ov);
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          o[e >> 1] = __floats2half2_rn(act_ct<ACT>(acc[j][e] * sc[e] + sh[e]), act_ct<ACT>(acc[j][e + 1] * sc[e + 1] + sh[e + 1]));
          const float2 rf = __half22float2(o[e >> 1]);
          psum[e] += rf.x; psum[e + 1] += rf.y;
        }
        *reinterpret_cast<uint4*>(p.out + ((long long)(n * p.Ho + oy) * p.Wo + ox) * p.out_cs + c0) = ov;
      }
    }
  }
  if (p.pool) {
#pragma unroll
    for (int e = 0; e < 8; ++e) s_pool[threadIdx.x * 8 + e] = psum[e];
    __syncthreads();
    if (threadIdx.x < kDw2Cb && cb0 + (int)threadIdx.x < p.C) {          // one thread per channel, fixed summation order
      const int vv = threadIdx.x >> 3, e = threadIdx.x & 7;
      float t = 0.0f;
      for (int q = 0; q < kDw2Threads / 4; ++q) t += s_pool[(q * 4 + vv) * 8 + e];
      p.pool[((long long)n * gridDim.x + blockIdx.x) * p.C + cb0 + threadIdx.x] = t;
    }
  }
}

// Split-fp16 ("strict" precision) depthwise conv: the tiled kernel's CTA geometry, pooling-partial layout and FFMA2 inner loop, with
// the window staged as fp32 (hi + lo summed once per element while staging: 144-byte pixel pitch = the low halves of the four
// 8-channel vectors, their high halves, 16 B of padding), fp32 filter taps and fp32 pooling of the exact (unrounded) outputs.
// Zero padding is staged as 0: fmaf(0, w, acc) == acc, so the result equals the tap-skipping form bit for bit.
constexpr int kDwsPitch = 144;
template <int K, int S, int XPT>
struct DwsCfg {
  using B = Dw2Cfg<K, S, XPT>;
  static constexpr int kInBytes = B::IW * B::IH * kDwsPitch;
  static constexpr int kWBytes = K * K * kDw2Cb * 4;
  static constexpr int kSmem = kInBytes + kWBytes + kDw2Threads * 8 * 4;
};

template <int K, int S, int XPT, int ACT>
__global__ void __launch_bounds__(kDw2Threads) depthwise_split_kernel(const DwParams p, int tiles_x, const float* __restrict__ w32, int in_lo, int out_lo) {
  using Cfg = Dw2Cfg<K, S, XPT>;
  extern __shared__ __align__(16) unsigned char dw_smem[];
  unsigned char* s_in = dw_smem;
  float* s_w = reinterpret_cast<float*>(dw_smem + DwsCfg<K, S, XPT>::kInBytes);
  float* s_pool = reinterpret_cast<float*>(dw_smem + DwsCfg<K, S, XPT>::kInBytes + DwsCfg<K, S, XPT>::kWBytes);
  const int n = blockIdx.z, cb0 = blockIdx.y * kDw2Cb;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const int oy0 = ty * Cfg::TOH, ox0 = tx * Cfg::TOW;
  const int iy0 = oy0 * S - p.pad, ix0 = ox0 * S - p.pad;
  const int nvec = min(4, (p.C - cb0) >> 3);
  const __half* inb = p.in + (long long)n * p.H * p.W * p.in_cs + cb0;
  // ---- stage the window as fp32: four (pixel, vector) elements per thread and pass, their eight 16-byte loads issued together
  constexpr int kElems = Cfg::IW * Cfg::IH * 4;
  for (int base = threadIdx.x; base < kElems; base += 4 * kDw2Threads) {
    uint4 hi[4], lo[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * kDw2Threads;
      const int v = i & 3, px = i >> 2;
      const int ly = px / Cfg::IW, lx = px - ly * Cfg::IW;
      const int iy = iy0 + ly, ix = ix0 + lx;
      hi[u] = make_uint4(0u, 0u, 0u, 0u); lo[u] = hi[u];
      if (i < kElems && v < nvec && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
        const __half* src = inb + ((long long)iy * p.W + ix) * p.in_cs + v * 8;
        hi[u] = __ldg(reinterpret_cast<const uint4*>(src));
        lo[u] = __ldg(reinterpret_cast<const uint4*>(src + in_lo));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * kDw2Threads;
      if (i >= kElems) continue;
      const int v = i & 3, px = i >> 2;
      const __half2* hh = reinterpret_cast<const __half2*>(&hi[u]);
      const __half2* lh = reinterpret_cast<const __half2*>(&lo[u]);
      float f[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 a = __half22float2(hh[e]), b = __half22float2(lh[e]);
        f[2 * e] = a.x + b.x; f[2 * e + 1] = a.y + b.y;
      }
      *reinterpret_cast<float4*>(s_in + px * kDwsPitch + v * 16) = make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4*>(s_in + px * kDwsPitch + 64 + v * 16) = make_float4(f[4], f[5], f[6], f[7]);
    }
  }
  for (int i = threadIdx.x; i < K * K * kDw2Cb; i += kDw2Threads) {
    const int t = i / kDw2Cb, c = i - t * kDw2Cb;
    s_w[i] = (cb0 + c < p.C) ? __ldg(w32 + (long long)t * p.C + cb0 + c) : 0.0f;
  }
  __syncthreads();
  const int v = threadIdx.x & 3, strip = (threadIdx.x >> 2) & 7, row = threadIdx.x >> 5;
  const int c0 = cb0 + v * 8;
  float2 acc2[XPT][4];
#pragma unroll
  for (int j = 0; j < XPT; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc2[j][e] = make_float2(0.0f, 0.0f);
  const unsigned char* base = s_in + ((row * S) * Cfg::IW + strip * XPT * S) * kDwsPitch + v * 16;
#pragma unroll
  for (int ky = 0; ky < K; ++ky) {
    float2 xf[Cfg::NX][4];
#pragma unroll
    for (int i = 0; i < Cfg::NX; ++i) {
      const float4 a = *reinterpret_cast<const float4*>(base + (ky * Cfg::IW + i) * kDwsPitch);
      const float4 b = *reinterpret_cast<const float4*>(base + (ky * Cfg::IW + i) * kDwsPitch + 64);
      xf[i][0] = make_float2(a.x, a.y); xf[i][1] = make_float2(a.z, a.w); xf[i][2] = make_float2(b.x, b.y); xf[i][3] = make_float2(b.z, b.w);
    }
#pragma unroll
    for (int kx = 0; kx < K; ++kx) {
      const float4 w0 = *reinterpret_cast<const float4*>(s_w + (ky * K + kx) * kDw2Cb + v * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(s_w + (ky * K + kx) * kDw2Cb + v * 8 + 4);
      const float2 wf[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y), make_float2(w1.z, w1.w)};
#pragma unroll
      for (int j = 0; j < XPT; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc2[j][e] = his_ffma2(xf[j * S + kx][e], wf[e], acc2[j][e]);
    }
  }
  float psum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int oy = oy0 + row;
  if (v < nvec && oy < p.Ho) {
    float sc[8], sh[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { sc[e] = __ldg(p.scale + c0 + e); sh[e] = __ldg(p.shift + c0 + e); }
#pragma unroll
    for (int j = 0; j < XPT; ++j) {
      const int ox = ox0 + strip * XPT + j;
      if (ox < p.Wo) {
        float y[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          y[2 * e] = act_ct<ACT>(acc2[j][e].x * sc[2 * e] + sh[2 * e]);
          y[2 * e + 1] = act_ct<ACT>(acc2[j][e].y * sc[2 * e + 1] + sh[2 * e + 1]);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) psum[e] += y[e];
        his_st8(p.out + ((long long)(n * p.Ho + oy) * p.Wo + ox) * p.out_cs + c0, out_lo, y);
      }
    }
  }
  if (p.pool) {
#pragma unroll
    for (int e = 0; e < 8; ++e) s_pool[threadIdx.x * 8 + e] = psum[e];
    __syncthreads();
    if (threadIdx.x < kDw2Cb && cb0 + (int)threadIdx.x < p.C) {
      const int vv = threadIdx.x >> 3, e = threadIdx.x & 7;
      float t = 0.0f;
      for (int q = 0; q < kDw2Threads / 4; ++q) t += s_pool[(q * 4 + vv) * 8 + e];
      p.pool[((long long)n * gridDim.x + blockIdx.x) * p.C + cb0 + threadIdx.x] = t;
    }
  }
}

// per-(n,c) sums of an NHWC fp16 tensor (ChannelAttentionModule's adaptive_avg_pool2d): per-block partials
// [N][gridDim.x][C], reduced in a fixed order (deterministic); blockDim.x is a multiple of C/8.
__global__ void pool_sum_kernel(const __half* __restrict__ in, int HW, int C, int cs, int lo, float* __restrict__ pool) {
  extern __shared__ float s_pool[];
  const int n = blockIdx.y, cgs = C / 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long step = (long long)gridDim.x * blockDim.x;
  const __half* img = in + (long long)n * HW * cs;
  // blockDim.x is a multiple of cgs (threads_multiple_of), so a thread keeps ONE channel group and walks pixels with a constant
  // stride: one division up front instead of a 64-bit divide + modulo per 16-byte load (same elements, same order as idx += step)
  const long long idx0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int cg0 = (int)(idx0 % cgs);
  const long long pix0 = idx0 / cgs, pstep = step / cgs;
  auto addr = [&](long long p) { return reinterpret_cast<const uint4*>(img + p * cs + cg0 * 8); };
  auto add = [&](const uint4& xv) {
    const __half2* xh = reinterpret_cast<const __half2*>(&xv);
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(xh[e]); acc[2 * e] += f.x; acc[2 * e + 1] += f.y; }
  };
  long long pix = pix0;
  for (; pix + 3 * pstep < HW; pix += 4 * pstep) {      // four loads in flight; the per-thread summation order is unchanged
    const uint4 v0 = __ldg(addr(pix)), v1 = __ldg(addr(pix + pstep)), v2 = __ldg(addr(pix + 2 * pstep)), v3 = __ldg(addr(pix + 3 * pstep));
    add(v0); add(v1); add(v2); add(v3);
  }
  for (; pix < HW; pix += pstep) add(__ldg(addr(pix)));
  if (lo)      // split-fp16 input: the lo planes of the same pixels (their sum is ~2^-11 of the total: order is immaterial)
    for (pix = pix0; pix < HW; pix += pstep) add(__ldg(addr(pix) + (lo >> 3)));
#pragma unroll
  for (int e = 0; e < 8; ++e) s_pool[threadIdx.x * 8 + e] = acc[e];
  __syncthreads();
  float* dst = pool + ((long long)n * gridDim.x + blockIdx.x) * C;
  const int first = (int)((blockIdx.x * (long long)blockDim.x) % cgs);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int cg = c >> 3, e = c & 7;
    float s = 0.0f;
    for (int t = (cg - first + cgs) % cgs; t < (int)blockDim.x; t += cgs) s += s_pool[t * 8 + e];
    dst[c] = s;
  }
}

// Sums the per-block pooling partials [N][nparts][C] into part 0.  Block = 32 channels x 32 part-lanes: lane y adds parts
// y, y+32, ... (coalesced over channels), then the 32 partial sums are added in a fixed order -> deterministic.  Each
// (n,c) of part 0 is written by exactly one thread after every read of the block (syncthreads), so in-place is race-free.
__global__ void pool_reduce_kernel(float* __restrict__ pool, int nparts, int C) {
  __shared__ float s[32][33];
  const int n = blockIdx.y, c = blockIdx.x * 32 + threadIdx.x;
  float* base = pool + (long long)n * nparts * C;
  float a = 0.0f;
  if (c < C)
    for (int part = threadIdx.y; part < nparts; part += 32) a += base[(long long)part * C + c];
  s[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.0f;
#pragma unroll
    for (int y = 0; y < 32; ++y) t += s[y][threadIdx.x];
    base[c] = t;
  }
}

// hidden[n][r] = act(W1[r,:] . (pool[n][part 0]/HW) + b1[r]): one warp per (n, r), lanes walk the contiguous weight row
// four elements at a time, fixed-order shuffle reduction.
__global__ void se_hidden_kernel(const float* __restrict__ pool, int nparts, float inv_hw, int C, int R, const float* __restrict__ w1,
                                 const float* __restrict__ b1, int act, float act_beta, float* __restrict__ hidden) {
  const int n = blockIdx.y, lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const float* mean = pool + (long long)n * nparts * C;
  const float* wr = w1 + (long long)r * C;
  float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
  int c = lane;
  for (; c + 96 < C; c += 128) {
    s0 = fmaf(__ldg(wr + c), mean[c], s0); s1 = fmaf(__ldg(wr + c + 32), mean[c + 32], s1);
    s2 = fmaf(__ldg(wr + c + 64), mean[c + 64], s2); s3 = fmaf(__ldg(wr + c + 96), mean[c + 96], s3);
  }
  for (; c < C; c += 32) s0 = fmaf(__ldg(wr + c), mean[c], s0);
  float s = (s0 + s1) + (s2 + s3);
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) hidden[(long long)n * R + r] = his_act(s * inv_hw + (b1 ? b1[r] : 0.0f), act, act_beta);
}

// gate[n][c] = sigmoid(W2[c,:] . hidden[n] + b2[c]): one warp per (n, c).
__global__ void se_gate_kernel(const float* __restrict__ hidden, int C, int R, const float* __restrict__ w2, const float* __restrict__ b2,
                               float* __restrict__ gate) {
  const int n = blockIdx.y, lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= C) return;
  const float* h = hidden + (long long)n * R;
  const float* wr = w2 + (long long)c * R;
  float s = 0.0f;
  for (int r = lane; r < R; r += 32) s = fmaf(__ldg(wr + r), h[r], s);
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) gate[(long long)n * C + c] = his_sigmoid(s + (b2 ? b2[c] : 0.0f));
}

// Folds a per-(image, input-channel) gate into the packed GEMM weights: wout[n][row][k] = w[row][k] * gate[n][k].
// conv(x * gate) == conv_{w*gate}(x), so the squeeze-excite product never touches the activation tensor; the projection
// GEMM then reads image n's weight slab (his_conv_gemm_set_image_weights).
__global__ void scale_weights_kernel(const __half* __restrict__ w, const float* __restrict__ gate, long long rows, int K, int C,
                                     long long total, __half* __restrict__ out) {
  const int kgs = K / 8;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int kg = (int)(idx % kgs);
    const long long row = (idx / kgs) % rows;
    const long long n = idx / (kgs * rows);
    uint4 wv = __ldg(reinterpret_cast<const uint4*>(w + row * K + kg * 8));
    __half2* wh = reinterpret_cast<__half2*>(&wv);
    const float* g = gate + n * C + kg * 8;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k0 = kg * 8 + 2 * e;
      const float2 f = __half22float2(wh[e]);
      wh[e] = __floats2half2_rn(k0 < C ? f.x * __ldg(g + 2 * e) : 0.0f, k0 + 1 < C ? f.y * __ldg(g + 2 * e + 1) : 0.0f);
    }
    *reinterpret_cast<uint4*>(out + (n * rows + row) * K + kg * 8) = wv;
  }
}

// split-fp16 form: rows are [W_hi | W_lo] (K1 elements each); out = split((W_hi + W_lo) * gate)
__global__ void scale_weights_split_kernel(const __half* __restrict__ w, const float* __restrict__ gate, long long rows, int K1, int C,
                                           long long total, __half* __restrict__ out) {
  const int kgs = K1 / 8;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int kg = (int)(idx % kgs);
    const long long row = (idx / kgs) % rows;
    const long long n = idx / (kgs * rows);
    float f[8];
    his_ld8(w + row * 2 * K1 + kg * 8, K1, f);
    const float* g = gate + n * C + kg * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = (kg * 8 + e < C) ? f[e] * __ldg(g + e) : 0.0f;
    his_st8(out + (n * rows + row) * 2 * K1 + kg * 8, K1, f);
  }
}

// x[n,p,c] *= gate[n,c]  (in place or to another slice)
__global__ void scale_channels_kernel(const __half* __restrict__ in, int in_cs, const float* __restrict__ gate, long long HW, int C,
                                      long long total, __half* __restrict__ out, int out_cs, int in_lo, int out_lo) {
  const int cgs = C / 8;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(idx % cgs);
    const long long pix = idx / cgs;
    const long long n = pix / HW;
    float f[8];
    his_ld8(in + pix * in_cs + cg * 8, in_lo, f);
    const float* g = gate + n * C + cg * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] *= __ldg(g + e);
    his_st8(out + pix * out_cs + cg * 8, out_lo, f);
  }
}

// ------------------------------------------------------------------------------------ LayerNorm2d (hed/model.py:18-38)
// Statistics over (C,H,W) per sample (biased variance, eps 1e-5), affine [C].  Two launches: fixed-order partial sums
// (fp32 per thread, double per block -> bit-reproducible), then normalise + affine (+residual) + activation.
__global__ void ln_stats_kernel(const __half* __restrict__ in, long long per_img_vec, int HW, int C, int cs, int lo, double* __restrict__ partials) {
  __shared__ double s_sum[kThreads], s_sq[kThreads];
  const int n = blockIdx.y, cgs = C / 8;
  float a = 0.0f, b = 0.0f;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < per_img_vec; idx += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(idx % cgs);
    const long long pix = idx / cgs;
    float f[8];
    his_ld8(in + ((long long)n * HW + pix) * cs + cg * 8, lo, f);
#pragma unroll
    for (int e = 0; e < 4; ++e) { a += f[2 * e] + f[2 * e + 1]; b = fmaf(f[2 * e], f[2 * e], fmaf(f[2 * e + 1], f[2 * e + 1], b)); }
  }
  s_sum[threadIdx.x] = (double)a; s_sq[threadIdx.x] = (double)b;
  __syncthreads();
  for (int o = kThreads / 2; o; o >>= 1) {
    if ((int)threadIdx.x < o) { s_sum[threadIdx.x] += s_sum[threadIdx.x + o]; s_sq[threadIdx.x] += s_sq[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partials[((long long)n * gridDim.x + blockIdx.x) * 2] = s_sum[0];
    partials[((long long)n * gridDim.x + blockIdx.x) * 2 + 1] = s_sq[0];
  }
}

__global__ void ln_apply_kernel(const __half* __restrict__ in, int HW, int C, int cs, const double* __restrict__ partials, int nparts,
                                const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int act, float act_beta,
                                int res_mode, const __half* __restrict__ res, int res_cs, __half* __restrict__ out, int out_cs,
                                int in_lo, int res_lo, int out_lo) {
  __shared__ float s_mu, s_rstd;
  const int n = blockIdx.y, cgs = C / 8;
  if (threadIdx.x < 32) {
    // fixed order whatever the batch: lane l adds parts l, l + 32, ... in sequence, then a butterfly over the lanes
    double s = 0.0, q = 0.0;
    for (int i = threadIdx.x; i < nparts; i += 32) { s += partials[((long long)n * nparts + i) * 2]; q += partials[((long long)n * nparts + i) * 2 + 1]; }
#pragma unroll
    for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if (threadIdx.x == 0) {
      const double cnt = (double)HW * (double)C;
      const double mu = s / cnt;
      double var = q / cnt - mu * mu;
      if (var < 0.0) var = 0.0;
      s_mu = (float)mu; s_rstd = (float)(1.0 / sqrt(var + (double)eps));
    }
  }
  __syncthreads();
  const float mu = s_mu, rstd = s_rstd;
  const long long per_img_vec = (long long)HW * cgs;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  if (nthreads % cgs == 0 && in_lo == 0 && out_lo == 0 && res_lo == 0) {
    // plain fp16 tensors: four pixels of the thread's channel group in flight as raw 16-byte vectors (the kernel is a pure HBM
    // stream; with ~70 registers three blocks fit an SM, so the bytes in flight have to come from the unroll)
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int cg = (int)(t % cgs);
    const int pstep = (int)(nthreads / cgs);
    float gm[8], bt[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { gm[e] = __ldg(gamma + cg * 8 + e); bt[e] = __ldg(beta + cg * 8 + e); }
    const __half* inp = in + (long long)n * HW * cs + cg * 8;
    const __half* resp = res_mode ? res + (long long)n * HW * res_cs + cg * 8 : nullptr;
    __half* outp = out + (long long)n * HW * out_cs + cg * 8;
    for (int pix = (int)(t / cgs); pix < HW; pix += 4 * pstep) {
      uint4 a[4], r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int pu = pix + u * pstep;
        a[u] = make_uint4(0u, 0u, 0u, 0u); r[u] = a[u];
        if (pu < HW) {
          a[u] = __ldg(reinterpret_cast<const uint4*>(inp + (long long)pu * cs));
          if (res_mode) r[u] = __ldg(reinterpret_cast<const uint4*>(resp + (long long)pu * res_cs));
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int pu = pix + u * pstep;
        if (pu < HW) {
          const __half2* ah = reinterpret_cast<const __half2*>(&a[u]);
          const __half2* rh = reinterpret_cast<const __half2*>(&r[u]);
          __half2 o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(ah[e]), rr = __half22float2(rh[e]);
            float y0 = (f.x - mu) * rstd * gm[2 * e] + bt[2 * e], y1 = (f.y - mu) * rstd * gm[2 * e + 1] + bt[2 * e + 1];
            if (res_mode == HIS_RES_ADD) { y0 += rr.x; y1 += rr.y; }
            y0 = his_act(y0, act, act_beta); y1 = his_act(y1, act, act_beta);
            if (res_mode == HIS_RES_MUL) { y0 *= rr.x; y1 *= rr.y; }
            o[e] = __floats2half2_rn(y0, y1);
          }
          *reinterpret_cast<uint4*>(outp + (long long)pu * out_cs) = *reinterpret_cast<const uint4*>(o);
        }
      }
    }
    return;
  }
  if (nthreads % cgs == 0) {
    // the grid stride is a multiple of the channel-vector count (the host sizes the grid so): a thread keeps ONE 8-channel group for
    // all its pixels -- gamma / beta live in registers, no 64-bit division per vector, two independent pixels in flight
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int cg = (int)(t % cgs);
    const int pstep = (int)(nthreads / cgs);
    float gm[8], bt[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { gm[e] = __ldg(gamma + cg * 8 + e); bt[e] = __ldg(beta + cg * 8 + e); }
    const __half* inp = in + (long long)n * HW * cs + cg * 8;
    const __half* resp = res_mode ? res + (long long)n * HW * res_cs + cg * 8 : nullptr;
    __half* outp = out + (long long)n * HW * out_cs + cg * 8;
    for (int pix = (int)(t / cgs); pix < HW; pix += 2 * pstep) {
      const int pix1 = pix + pstep;
      const bool two = pix1 < HW;
      float f0[8], f1[8], r0[8] = {0, 0, 0, 0, 0, 0, 0, 0}, r1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      his_ld8(inp + (long long)pix * cs, in_lo, f0);
      if (two) his_ld8(inp + (long long)pix1 * cs, in_lo, f1);
      if (res_mode) {
        his_ld8(resp + (long long)pix * res_cs, res_lo, r0);
        if (two) his_ld8(resp + (long long)pix1 * res_cs, res_lo, r1);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float y = (f0[e] - mu) * rstd * gm[e] + bt[e];
        if (res_mode == HIS_RES_ADD) y += r0[e];
        y = his_act(y, act, act_beta);
        if (res_mode == HIS_RES_MUL) y *= r0[e];
        f0[e] = y;
      }
      his_st8(outp + (long long)pix * out_cs, out_lo, f0);
      if (two) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float y = (f1[e] - mu) * rstd * gm[e] + bt[e];
          if (res_mode == HIS_RES_ADD) y += r1[e];
          y = his_act(y, act, act_beta);
          if (res_mode == HIS_RES_MUL) y *= r1[e];
          f1[e] = y;
        }
        his_st8(outp + (long long)pix1 * out_cs, out_lo, f1);
      }
    }
    return;
  }
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < per_img_vec; idx += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(idx % cgs);
    const long long pix = (long long)n * HW + idx / cgs;
    float f[8], r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    his_ld8(in + pix * cs + cg * 8, in_lo, f);
    if (res_mode) his_ld8(res + pix * res_cs + cg * 8, res_lo, r);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = cg * 8 + e;
      float y = (f[e] - mu) * rstd * __ldg(gamma + c) + __ldg(beta + c);
      if (res_mode == HIS_RES_ADD) y += r[e];
      y = his_act(y, act, act_beta);
      if (res_mode == HIS_RES_MUL) y *= r[e];
      f[e] = y;
    }
    his_st8(out + pix * out_cs + cg * 8, out_lo, f);
  }
}

// ------------------------------------------------------------------------------------ GroupNorm family
// nn.GroupNorm / nn.InstanceNorm2d(affine) / AdaptiveInstanceNorm2d / SpatialGroupNorm of get_normalization_layer
// (hed/advanced/normalization_comparison.py:12-74,159-206): statistics per (sample, group of C/G channels) over (C/G, H, W), biased
// variance, affine [C].  Three launches, fixed summation order (bit-reproducible):
//   1. per-channel partial (sum, sum of squares) of a pixel range: thread = (8-channel vector, pixel lane), fp32 per thread,
//      pixel lanes merged in order through shared memory            -> ws[n][part][c][2]
//   2. per (n, group): parts and channels summed in double          -> ws[n][parts][c] = (mean, rstd) of the channel's group
//   3. normalise + affine (+ residual) + activation, 16-byte vectors
constexpr int kGnMaxC = 2048;
__global__ void gn_stats_kernel(const __half* __restrict__ in, int HW, int C, int cs, int lo, int pix_per_part, int nslots, float* __restrict__ ws) {
  extern __shared__ float s_gn[];                       // [pixel lanes][C][2]
  const int n = blockIdx.y, part = blockIdx.x, cgs = C / 8;
  const int PL = blockDim.x / cgs;                      // pixel lanes (>= 1: C <= 2048)
  const int cg = threadIdx.x % cgs, pl = threadIdx.x / cgs;
  const int p0 = part * pix_per_part, p1 = min(p0 + pix_per_part, HW);
  float sum[8], sq[8], pivot[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { sum[e] = 0.0f; sq[e] = 0.0f; }
  if (pl < PL) {
    // sums of (x - pivot) and (x - pivot)^2 with pivot = the channel's value at the image's first pixel: a channel whose deviation
    // is small against its mean (instance norm on near-constant maps) would lose E[x^2] - E[x]^2 to cancellation in fp32
    his_ld8(in + (long long)n * HW * cs + cg * 8, lo, pivot);
    for (int pix = p0 + pl; pix < p1; pix += PL) {
      float f[8];
      his_ld8(in + ((long long)n * HW + pix) * cs + cg * 8, lo, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float d = f[e] - pivot[e]; sum[e] += d; sq[e] = fmaf(d, d, sq[e]); }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) { s_gn[((size_t)pl * C + cg * 8 + e) * 2] = sum[e]; s_gn[((size_t)pl * C + cg * 8 + e) * 2 + 1] = sq[e]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.0f, b = 0.0f;
    for (int l = 0; l < PL; ++l) { a += s_gn[((size_t)l * C + c) * 2]; b += s_gn[((size_t)l * C + c) * 2 + 1]; }
    float* dst = ws + (((long long)n * nslots + part) * C + c) * 2;
    dst[0] = a; dst[1] = b;
  }
}

__global__ void gn_finalize_kernel(float* __restrict__ ws, const __half* __restrict__ in, int cs, int lo, int HW, int C, int G, int nparts, float eps) {
  const int n = blockIdx.y, g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const int gs = C / G;
  // per channel: mean_c = pivot_c + S1/HW and the squared deviations about it M2_c = S2 - S1^2/HW (both well conditioned), then the
  // group: mean_g = avg mean_c, M2_g = sum_c [M2_c + HW (mean_c - mean_g)^2]  (parallel-variance merge, in double)
  double mean_sum = 0.0;
  for (int c = 0; c < gs; ++c) {
    double s1 = 0.0;
    for (int part = 0; part < nparts; ++part) s1 += (double)ws[(((long long)n * (nparts + 1) + part) * C + g * gs + c) * 2];
    mean_sum += (double)his_ld1(in + (long long)n * HW * cs + g * gs + c, lo) + s1 / (double)HW;
  }
  const double mu = mean_sum / (double)gs;
  double m2 = 0.0;
  for (int c = 0; c < gs; ++c) {
    double s1 = 0.0, s2 = 0.0;
    for (int part = 0; part < nparts; ++part) {
      const float* src = ws + (((long long)n * (nparts + 1) + part) * C + g * gs + c) * 2;
      s1 += (double)src[0]; s2 += (double)src[1];
    }
    const double mc = (double)his_ld1(in + (long long)n * HW * cs + g * gs + c, lo) + s1 / (double)HW;
    m2 += (s2 - s1 * s1 / (double)HW) + (double)HW * (mc - mu) * (mc - mu);
  }
  const double cnt = (double)HW * (double)gs;
  double var = m2 / cnt;
  if (var < 0.0) var = 0.0;
  const float muf = (float)mu, rstd = (float)(1.0 / sqrt(var + (double)eps));
  float* dst = ws + (((long long)n * (nparts + 1) + nparts) * C + g * gs) * 2;
  for (int c = 0; c < gs; ++c) { dst[2 * c] = muf; dst[2 * c + 1] = rstd; }
}

__global__ void gn_apply_kernel(const __half* __restrict__ in, int HW, int C, int cs, const float* __restrict__ ws, int nparts,
                                const float* __restrict__ gamma, const float* __restrict__ beta, int act, float act_beta,
                                int res_mode, const __half* __restrict__ res, int res_cs, __half* __restrict__ out, int out_cs,
                                int in_lo, int res_lo, int out_lo) {
  extern __shared__ float s_ab[];                       // per channel: y = (x - mean)*a + beta with a = rstd*gamma (the subtraction
                                                        // first: x*a - mean*a would cancel for channels with |mean| >> deviation)
  const int n = blockIdx.y, cgs = C / 8;
  const float* st = ws + ((long long)n * (nparts + 1) + nparts) * C * 2;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    s_ab[3 * c] = st[2 * c]; s_ab[3 * c + 1] = st[2 * c + 1] * __ldg(gamma + c); s_ab[3 * c + 2] = __ldg(beta + c);
  }
  __syncthreads();
  const long long per_img_vec = (long long)HW * cgs;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < per_img_vec; idx += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(idx % cgs);
    const long long pix = (long long)n * HW + idx / cgs;
    float f[8], r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    his_ld8(in + pix * cs + cg * 8, in_lo, f);
    if (res_mode) his_ld8(res + pix * res_cs + cg * 8, res_lo, r);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = cg * 8 + e;
      float y = fmaf(f[e] - s_ab[3 * c], s_ab[3 * c + 1], s_ab[3 * c + 2]);
      if (res_mode == HIS_RES_ADD) y += r[e];
      y = his_act(y, act, act_beta);
      if (res_mode == HIS_RES_MUL) y *= r[e];
      f[e] = y;
    }
    his_st8(out + pix * out_cs + cg * 8, out_lo, f);
  }
}

// ForegroundAwareNorm (normalization_comparison.py:113-132): y = IN(x) * (p*fs + (1-p)*bs) + (p*fb + (1-p)*bb) with the per-pixel
// foreground probability p (fp32 [N*HW], from the detector convs) and the instance statistics of gn_stats / gn_finalize (G = C).
__global__ void fgaware_apply_kernel(const __half* __restrict__ in, int HW, int C, int cs, const float* __restrict__ ws, int nparts,
                                     const float* __restrict__ fg_scale, const float* __restrict__ fg_bias, const float* __restrict__ bg_scale,
                                     const float* __restrict__ bg_bias, const float* __restrict__ prob, int act, float act_beta, int res_mode,
                                     const __half* __restrict__ res, int res_cs, __half* __restrict__ out, int out_cs, int in_lo, int res_lo,
                                     int out_lo) {
  extern __shared__ float s_fg[];                       // per channel: mean, rstd, fs, fb, bs, bb
  const int n = blockIdx.y, cgs = C / 8;
  const float* st = ws + ((long long)n * (nparts + 1) + nparts) * C * 2;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    s_fg[6 * c] = st[2 * c]; s_fg[6 * c + 1] = st[2 * c + 1];
    s_fg[6 * c + 2] = __ldg(fg_scale + c); s_fg[6 * c + 3] = __ldg(fg_bias + c);
    s_fg[6 * c + 4] = __ldg(bg_scale + c); s_fg[6 * c + 5] = __ldg(bg_bias + c);
  }
  __syncthreads();
  const long long per_img_vec = (long long)HW * cgs;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < per_img_vec; idx += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(idx % cgs);
    const long long pix = (long long)n * HW + idx / cgs;
    const float pf = __ldg(prob + pix), pb = 1.0f - pf;
    float f[8], r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    his_ld8(in + pix * cs + cg * 8, in_lo, f);
    if (res_mode) his_ld8(res + pix * res_cs + cg * 8, res_lo, r);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float* q = s_fg + 6 * (cg * 8 + e);
      const float xn = (f[e] - q[0]) * q[1];
      const float scale = pf * q[2] + pb * q[4], bias = pf * q[3] + pb * q[5];
      float y = xn * scale + bias;
      if (res_mode == HIS_RES_ADD) y += r[e];
      y = his_act(y, act, act_beta);
      if (res_mode == HIS_RES_MUL) y *= r[e];
      f[e] = y;
    }
    his_st8(out + pix * out_cs + cg * 8, out_lo, f);
  }
}

// ConvTranspose2d(k2,s2) for tiny Cin (upsample_bg_fg.0 in LayerNorm mode: 2 -> 32): NCHW fp32 in, NHWC fp16 out (+bias)
__global__ void convT2x2_small_kernel(const float* __restrict__ in, int N, int Cin, int h, int w, const float* __restrict__ wt,
                                      const float* __restrict__ bias, int Cout, __half* __restrict__ out, int out_cs, int out_lo) {
  const int Ho = 2 * h, Wo = 2 * w;
  const long long total = (long long)N * Ho * Wo * Cout;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % Cout);
    const long long pix = idx / Cout;
    const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho), n = (int)(pix / ((long long)Wo * Ho));
    float v = bias ? bias[c] : 0.0f;
    for (int ci = 0; ci < Cin; ++ci)
      v = fmaf(in[((long long)(n * Cin + ci) * h + (oy >> 1)) * w + (ox >> 1)], wt[((ci * Cout + c) * 2 + (oy & 1)) * 2 + (ox & 1)], v);
    his_st1(out + pix * out_cs + c, out_lo, v);
  }
}

// ------------------------------------------------------------------------------------ pooling / resize
__global__ void maxpool2_kernel(const __half* __restrict__ in, int N, int H, int W, int C, int in_cs, __half* __restrict__ out, int out_cs,
                                int in_lo, int out_lo) {
  const int Ho = H / 2, Wo = W / 2, cgs = C / 8;
  const long long total = (long long)N * Ho * Wo * cgs;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(idx % cgs);
    const long long pix = idx / cgs;
    const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho), n = (int)(pix / ((long long)Wo * Ho));
    const __half* base = in + ((long long)(n * H + 2 * oy) * W + 2 * ox) * in_cs + cg * 8;
    if (in_lo) {       // split-fp16: the maximum of the hi + lo sums, re-split (exact: the winner's own pair)
      float fa[8], fb[8], fc[8], fd[8];
      his_ld8(base, in_lo, fa); his_ld8(base + in_cs, in_lo, fb);
      his_ld8(base + (long long)W * in_cs, in_lo, fc); his_ld8(base + (long long)(W + 1) * in_cs, in_lo, fd);
#pragma unroll
      for (int e = 0; e < 8; ++e) fa[e] = fmaxf(fmaxf(fa[e], fb[e]), fmaxf(fc[e], fd[e]));
      his_st8(out + pix * out_cs + cg * 8, out_lo, fa);
      continue;
    }
    uint4 a = __ldg(reinterpret_cast<const uint4*>(base)), b = __ldg(reinterpret_cast<const uint4*>(base + in_cs));
    uint4 c = __ldg(reinterpret_cast<const uint4*>(base + (long long)W * in_cs)), d = __ldg(reinterpret_cast<const uint4*>(base + (long long)(W + 1) * in_cs));
    __half2* ah = reinterpret_cast<__half2*>(&a); const __half2* bh = reinterpret_cast<const __half2*>(&b);
    const __half2* ch = reinterpret_cast<const __half2*>(&c); const __half2* dh = reinterpret_cast<const __half2*>(&d);
#pragma unroll
    for (int e = 0; e < 4; ++e) ah[e] = __hmax2(__hmax2(ah[e], bh[e]), __hmax2(ch[e], dh[e]));
    *reinterpret_cast<uint4*>(out + pix * out_cs + cg * 8) = a;
  }
}

// nearest resize of an NHWC fp16 tensor into a channel slice (smp UnetDecoderBlock F.interpolate(mode="nearest"))
__global__ void resize_nearest_kernel(const __half* __restrict__ in, int N, int H, int W, int C, int in_cs, int Ho, int Wo,
                                      __half* __restrict__ out, int out_cs, int in_lo, int out_lo) {
  const int cgs = C / 8;
  const float sy = (float)H / (float)Ho, sx = (float)W / (float)Wo;
  const long long total = (long long)N * Ho * Wo * cgs;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(idx % cgs);
    const long long pix = idx / cgs;
    const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho), n = (int)(pix / ((long long)Wo * Ho));
    const int iy = min((int)floorf(oy * sy), H - 1), ix = min((int)floorf(ox * sx), W - 1);
    *reinterpret_cast<uint4*>(out + pix * out_cs + cg * 8) =
        __ldg(reinterpret_cast<const uint4*>(in + ((long long)(n * H + iy) * W + ix) * in_cs + cg * 8));
    if (in_lo)
      *reinterpret_cast<uint4*>(out + pix * out_cs + out_lo + cg * 8) =
          __ldg(reinterpret_cast<const uint4*>(in + ((long long)(n * H + iy) * W + ix) * in_cs + in_lo + cg * 8));
  }
}

// bilinear resize (align_corners=False, PyTorch upsample_bilinear2d source-index rule) of NCHW fp32
__global__ void resize_bilinear_f32_kernel(const float* __restrict__ in, int NC, int H, int W, int Ho, int Wo, float* __restrict__ out) {
  const float sy = (float)H / (float)Ho, sx = (float)W / (float)Wo;
  const long long total = (long long)NC * Ho * Wo;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % Wo), oy = (int)((idx / Wo) % Ho);
    const long long nc = idx / ((long long)Wo * Ho);
    const float fy = fmaxf(((float)oy + 0.5f) * sy - 0.5f, 0.0f), fx = fmaxf(((float)ox + 0.5f) * sx - 0.5f, 0.0f);
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float* b = in + nc * H * W;
    out[idx] = (1.0f - ly) * ((1.0f - lx) * b[y0 * W + x0] + lx * b[y0 * W + x1]) + ly * ((1.0f - lx) * b[y1 * W + x0] + lx * b[y1 * W + x1]);
  }
}

// bilinear resize (align_corners=False) of an NHWC fp16 slice into another slice (MultiScaleRGBSegmentationModel resizes every
// scale's features to 28x28, rgb.py:887-893); 8 channels per thread, fp32 interpolation
__global__ void resize_bilinear_half_kernel(const __half* __restrict__ in, int N, int H, int W, int C, int in_cs, int Ho, int Wo,
                                            __half* __restrict__ out, int out_cs, int in_lo, int out_lo) {
  const int cgs = C / 8;
  const float sy = (float)H / (float)Ho, sx = (float)W / (float)Wo;
  const long long total = (long long)N * Ho * Wo * cgs;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(idx % cgs);
    const long long pix = idx / cgs;
    const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho), n = (int)(pix / ((long long)Wo * Ho));
    const float fy = fmaxf(((float)oy + 0.5f) * sy - 0.5f, 0.0f), fx = fmaxf(((float)ox + 0.5f) * sx - 0.5f, 0.0f);
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const __half* b = in + (long long)n * H * W * in_cs + cg * 8;
    float p00[8], p01[8], p10[8], p11[8], r[8];
    his_ld8(b + ((long long)y0 * W + x0) * in_cs, in_lo, p00); his_ld8(b + ((long long)y0 * W + x1) * in_cs, in_lo, p01);
    his_ld8(b + ((long long)y1 * W + x0) * in_cs, in_lo, p10); his_ld8(b + ((long long)y1 * W + x1) * in_cs, in_lo, p11);
#pragma unroll
    for (int e = 0; e < 8; ++e) r[e] = (1.0f - ly) * ((1.0f - lx) * p00[e] + lx * p01[e]) + ly * ((1.0f - lx) * p10[e] + lx * p11[e]);
    his_st8(out + pix * out_cs + cg * 8, out_lo, r);
  }
}

// ------------------------------------------------------------------------------------ attention glue
// SpatialAttentionModule (attention_modules.py:67-113): per-pixel channel mean & max -> [N,H,W,2] fp32
__global__ void channel_stats_kernel(const __half* __restrict__ in, long long pixels, int C, int cs, int lo, float* __restrict__ stats) {
  // one warp per pixel: lanes stride the channel vector in 16-byte pieces
  const int lane = threadIdx.x & 31;
  const long long warp_global = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long pix = warp_global; pix < pixels; pix += nwarps) {
    float s = 0.0f, m = -INFINITY;
    for (int c = lane * 8; c < C; c += 256) {
      float f[8];
      his_ld8(in + pix * cs + c, lo, f);
#pragma unroll
      for (int e = 0; e < 4; ++e) { s += f[2 * e] + f[2 * e + 1]; m = fmaxf(m, fmaxf(f[2 * e], f[2 * e + 1])); }
    }
    for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o)); }
    if (lane == 0) { stats[pix * 2] = s / (float)C; stats[pix * 2 + 1] = m; }
  }
}

// x * sigmoid(conv_kxk([mean,max])) ; w is [2][k][k] fp32 (PyTorch [1,2,k,k])
__global__ void spatial_attention_apply_kernel(const __half* __restrict__ in, int N, int H, int W, int C, int in_cs,
                                               const float* __restrict__ stats, const float* __restrict__ w, int k,
                                               __half* __restrict__ out, int out_cs, int in_lo, int out_lo) {
  const int lane = threadIdx.x & 31;
  const long long pixels = (long long)N * H * W;
  const long long warp_global = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int pad = k / 2;
  for (long long pix = warp_global; pix < pixels; pix += nwarps) {
    const int x = (int)(pix % W), y = (int)((pix / W) % H);
    const long long img0 = pix - ((long long)y * W + x);
    float s = 0.0f;
    for (int t = lane; t < k * k; t += 32) {
      const int ky = t / k, kx = t % k, iy = y + ky - pad, ix = x + kx - pad;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
        const float2 st = __ldg(reinterpret_cast<const float2*>(stats + (img0 + (long long)iy * W + ix) * 2));
        s = fmaf(st.x, __ldg(w + t), s); s = fmaf(st.y, __ldg(w + k * k + t), s);
      }
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float g = his_sigmoid(s);
    for (int c = lane * 8; c < C; c += 256) {
      float f[8];
      his_ld8(in + pix * in_cs + c, in_lo, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] *= g;
      his_st8(out + pix * out_cs + c, out_lo, f);
    }
  }
}

// sigmoid(conv_kxk([mean,max])) alone: the statistics come from the producing GEMM's epilogue and the gate is applied as a
// row scale inside the consuming GEMM (his_conv_gemm_set_row_ops), so the 256-channel tensor is not touched here
__global__ void spatial_gate_kernel(const float* __restrict__ stats, int N, int H, int W, const float* __restrict__ w, int k, float* __restrict__ gate) {
  const long long total = (long long)N * H * W;
  const int pad = k / 2;
  for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < total; pix += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(pix % W), y = (int)((pix / W) % H);
    const long long img0 = pix - ((long long)y * W + x);
    float s = 0.0f;
    for (int ky = 0; ky < k; ++ky) {
      const int iy = y + ky - pad;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int ix = x + kx - pad;
        if (ix < 0 || ix >= W) continue;
        const float2 st = __ldg(reinterpret_cast<const float2*>(stats + (img0 + (long long)iy * W + ix) * 2));
        s = fmaf(st.x, __ldg(w + ky * k + kx), s); s = fmaf(st.y, __ldg(w + k * k + ky * k + kx), s);
      }
    }
    gate[pix] = his_sigmoid(s);
  }
}

// ------------------------------------------------------------------------------------ head tails
// upsample_bg_fg (..._refinement.py:501-506): ConvT(2->32,k2,s2) + norm(BN folded) + act + 1x1(32->2), NCHW fp32 in/out.
//   wt: [2][32][2][2] (PyTorch ConvTranspose2d weight), s/t: folded scale/shift [32] (bias folded in), w1: [2][32], b1: [2]
// One thread = one INPUT pixel = its 2x2 output block, so every parameter is the same for all lanes: the folded taps
// A[c][ph] = wt[0][c][ph]*s[c], B[c][ph] = wt[1][c][ph]*s[c] (ph = ky*2+kx), t[c] and the 1x1 rows are broadcast reads from
// shared memory (the per-output-pixel form did six lane-divergent global loads per channel: 0.36 ms per B0 step).
__global__ void __launch_bounds__(kThreads) upsample_bgfg_kernel(const float* __restrict__ low, int N, int h, int w, const float* __restrict__ wt,
                                     const float* __restrict__ s, const float* __restrict__ t, const float* __restrict__ w1,
                                     const float* __restrict__ b1, int act, float act_beta, float* __restrict__ out) {
  __shared__ __align__(16) float s_a[32 * 4], s_b[32 * 4], s_t[32], s_w[2 * 32];
  for (int i = threadIdx.x; i < 128; i += blockDim.x) {
    const int c = i >> 2, ph = i & 3;
    s_a[i] = __ldg(wt + (0 * 32 + c) * 4 + ph) * __ldg(s + c);
    s_b[i] = __ldg(wt + (1 * 32 + c) * 4 + ph) * __ldg(s + c);
  }
  for (int i = threadIdx.x; i < 32; i += blockDim.x) s_t[i] = __ldg(t + i);
  for (int i = threadIdx.x; i < 64; i += blockDim.x) s_w[i] = __ldg(w1 + i);
  __syncthreads();
  const int Ho = 2 * h, Wo = 2 * w;
  const long long total = (long long)N * h * w;
  const float bias0 = __ldg(b1), bias1 = __ldg(b1 + 1);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int ix = (int)(idx % w), iy = (int)((idx / w) % h), n = (int)(idx / ((long long)w * h));
    const float a0 = low[((long long)(n * 2 + 0) * h + iy) * w + ix], a1 = low[((long long)(n * 2 + 1) * h + iy) * w + ix];
    float o0[4] = {bias0, bias0, bias0, bias0}, o1[4] = {bias1, bias1, bias1, bias1};
#pragma unroll 8
    for (int c = 0; c < 32; ++c) {
      const float4 A = *reinterpret_cast<const float4*>(s_a + 4 * c), B = *reinterpret_cast<const float4*>(s_b + 4 * c);
      const float tc = s_t[c], u0 = s_w[c], u1 = s_w[32 + c];
      const float v[4] = {his_act(fmaf(a0, A.x, fmaf(a1, B.x, tc)), act, act_beta), his_act(fmaf(a0, A.y, fmaf(a1, B.y, tc)), act, act_beta),
                          his_act(fmaf(a0, A.z, fmaf(a1, B.z, tc)), act, act_beta), his_act(fmaf(a0, A.w, fmaf(a1, B.w, tc)), act, act_beta)};
#pragma unroll
      for (int ph = 0; ph < 4; ++ph) { o0[ph] = fmaf(v[ph], u0, o0[ph]); o1[ph] = fmaf(v[ph], u1, o1[ph]); }
    }
    float* d0 = out + ((long long)(n * 2 + 0) * Ho + 2 * iy) * Wo + 2 * ix;
    float* d1 = out + ((long long)(n * 2 + 1) * Ho + 2 * iy) * Wo + 2 * ix;
    *reinterpret_cast<float2*>(d0) = make_float2(o0[0], o0[1]); *reinterpret_cast<float2*>(d0 + Wo) = make_float2(o0[2], o0[3]);
    *reinterpret_cast<float2*>(d1) = make_float2(o1[0], o1[1]); *reinterpret_cast<float2*>(d1 + Wo) = make_float2(o1[2], o1[3]);
  }
}

// ..._refinement.py:588-596: L0 = bgfg0, L1 = bgfg1 + tn0*softmax(bgfg)1, L2 = bgfg1 + tn1*softmax(bgfg)1  (NCHW fp32)
__global__ void head_combine_kernel(const float* __restrict__ bgfg, const float* __restrict__ tn, int N, long long HW, float* __restrict__ logits) {
  const long long total = (long long)N * HW;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long n = idx / HW, p = idx % HW;
    const float b0 = bgfg[(n * 2) * HW + p], b1 = bgfg[(n * 2 + 1) * HW + p];
    const float t0 = tn[(n * 2) * HW + p], t1 = tn[(n * 2 + 1) * HW + p];
    const float m = fmaxf(b0, b1);
    const float e0 = expf(b0 - m), e1 = expf(b1 - m);
    const float fg = e1 / (e0 + e1);
    logits[(n * 3) * HW + p] = b0;
    logits[(n * 3 + 1) * HW + p] = b1 + t0 * fg;
    logits[(n * 3 + 2) * HW + p] = b1 + t1 * fg;
  }
}

// elementwise map on fp32: 0 sigmoid(x) ; 1 sigmoid((x - *param) * 10)  (DistanceTransformDecoder ..._refinement.py:341)
__global__ void map_f32_kernel(const float* __restrict__ in, long long total, int op, const float* __restrict__ param, float* __restrict__ out) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const float x = in[idx];
    out[idx] = op == 0 ? 1.0f / (1.0f + expf(-x)) : 1.0f / (1.0f + expf(-((x - __ldg(param)) * 10.0f)));
  }
}

// ---- PretrainedUNetGuidedSegmentationHead glue (rgb.py:125-218)
// fg_prob = sigmoid(channel c of an NCHW fp32 tensor) -> channel 0 of an NHWC fp16 slice (the concat slot) and/or fp32 [N,HW]
__global__ void sigmoid_channel_kernel(const float* __restrict__ in, int C, long long HW, int c, long long total, __half* __restrict__ out_h,
                                       int out_cs, int out_lo, float* __restrict__ out_f) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long n = idx / HW, p = idx - n * HW;
    const float v = 1.0f / (1.0f + expf(-in[(n * C + c) * HW + p]));
    if (out_h) his_st1(out_h + idx * out_cs, out_lo, v);
    if (out_f) out_f[idx] = v;
  }
}

// processed * (attention * (0.5 + 0.5 * fg_prob))  (rgb.py:169-173); a, f: fp32 per pixel
__global__ void scale_pixels_kernel(const __half* __restrict__ in, int in_cs, const float* __restrict__ a, const float* __restrict__ f, int C,
                                    long long total, __half* __restrict__ out, int out_cs, int in_lo, int out_lo) {
  const int cgs = C / 8;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(idx % cgs);
    const long long pix = idx / cgs;
    const float g = __ldg(a + pix) * (0.5f + 0.5f * __ldg(f + pix));
    float v[8];
    his_ld8(in + pix * in_cs + cg * 8, in_lo, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] *= g;
    his_st8(out + pix * out_cs + cg * 8, out_lo, v);
  }
}

// aux outputs of the guided head (rgb.py:187-216): m = bilinear(mask channel c -> (Ho,Wo)) (identity when equal),
// fg = sigmoid(m), bg_fg_logits = [log(1-fg+1e-7), log(fg+1e-7)]
__global__ void guided_aux_kernel(const float* __restrict__ in, int C, int c, int H, int W, int Ho, int Wo, long long total,
                                  float* __restrict__ mask_out, float* __restrict__ fg_out, float* __restrict__ bgfg_out) {
  const float sy = (float)H / (float)Ho, sx = (float)W / (float)Wo;
  const bool same = H == Ho && W == Wo;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % Wo), oy = (int)((idx / Wo) % Ho);
    const long long n = idx / ((long long)Wo * Ho);
    const float* b = in + (n * C + c) * (long long)H * W;
    float m;
    if (same) m = b[oy * W + ox];
    else {
      const float fy = fmaxf(((float)oy + 0.5f) * sy - 0.5f, 0.0f), fx = fmaxf(((float)ox + 0.5f) * sx - 0.5f, 0.0f);
      const int y0 = (int)fy, x0 = (int)fx;
      const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
      const float ly = fy - (float)y0, lx = fx - (float)x0;
      m = (1.0f - ly) * ((1.0f - lx) * b[y0 * W + x0] + lx * b[y0 * W + x1]) + ly * ((1.0f - lx) * b[y1 * W + x0] + lx * b[y1 * W + x1]);
    }
    const float fg = 1.0f / (1.0f + expf(-m));
    const long long hw = (long long)Ho * Wo, p = idx - n * hw;
    mask_out[idx] = m;
    fg_out[idx] = fg;
    bgfg_out[(n * 2) * hw + p] = logf((1.0f - fg) + 1e-7f);
    bgfg_out[(n * 2 + 1) * hw + p] = logf(fg + 1e-7f);
  }
}

// depth-to-space of an NHWC fp16 slice: out[n][2y+py][2x+px][c] = in[n][y][x][(py*2+px)*C + c].  ConvTranspose2d(k4, s2, p1)
// of ProgressiveUpsamplingDecoder (..._refinement.py:152-215) is four 2x2 phase convolutions of the input; they run as ONE 3x3
// conv with 4*C output channels (phase-major, the taps a phase does not use are zero) and this kernel interleaves the phases.
__global__ void depth_to_space2_half_kernel(const __half* __restrict__ in, int N, int h, int w, int C, int in_cs, __half* __restrict__ out, int out_cs,
                                            int in_lo, int out_lo) {
  const int cgs = C / 8, Ho = 2 * h, Wo = 2 * w;
  const long long total = (long long)N * Ho * Wo * cgs;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(idx % cgs);
    const long long pix = idx / cgs;
    const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho), n = (int)(pix / ((long long)Wo * Ho));
    const int ph = (oy & 1) * 2 + (ox & 1);
    *reinterpret_cast<uint4*>(out + pix * out_cs + cg * 8) =
        __ldg(reinterpret_cast<const uint4*>(in + ((long long)(n * h + (oy >> 1)) * w + (ox >> 1)) * in_cs + ph * C + cg * 8));
    if (in_lo)
      *reinterpret_cast<uint4*>(out + pix * out_cs + out_lo + cg * 8) =
          __ldg(reinterpret_cast<const uint4*>(in + ((long long)(n * h + (oy >> 1)) * w + (ox >> 1)) * in_cs + in_lo + ph * C + cg * 8));
  }
}

// ---- refinement flags of the refined head (hierarchical_segmentation_refinement.py)
// PixelShuffle(2) of NCHW fp32 (SubPixelDecoder :218-252): out[n][c][2y+i][2x+j] = in[n][c*4 + i*2 + j][y][x]; the input holds
// in_ch >= 4*C channels per image (the producing GEMM pads 12 to 16)
__global__ void pixel_shuffle2_kernel(const float* __restrict__ in, int N, int C, int in_ch, int h, int w, float* __restrict__ out) {
  const int Ho = 2 * h, Wo = 2 * w;
  const long long total = (long long)N * C * Ho * Wo;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % Wo), oy = (int)((idx / Wo) % Ho);
    const long long nc = idx / ((long long)Wo * Ho);
    const long long n = nc / C; const int c = (int)(nc - n * C);
    out[idx] = in[((n * in_ch + c * 4 + (oy & 1) * 2 + (ox & 1)) * h + (oy >> 1)) * (long long)w + (ox >> 1)];
  }
}

__device__ __forceinline__ void softmax3(const float* __restrict__ l, long long HW, long long p, float* pr) {
  const float a = l[p], b = l[HW + p], c = l[2 * HW + p];
  const float m = fmaxf(a, fmaxf(b, c));
  const float ea = expf(a - m), eb = expf(b - m), ec = expf(c - m);
  const float s = (ea + eb) + ec;
  pr[0] = ea / s; pr[1] = eb / s; pr[2] = ec / s;
}

// BoundaryRefinementModule.detect_edges (:94-129) before the normalisation: mean_c sqrt(dy^2 + dx^2) of the softmax gradients
// (forward differences, last row / column replicated), and the min / max over the WHOLE batch tensor (non-negative floats order
// like their bit patterns -> integer atomics, order independent).
__global__ void boundary_edges_kernel(const float* __restrict__ logits, int N, int H, int W, float* __restrict__ edges, unsigned int* __restrict__ minmax) {
  const long long HW = (long long)H * W, total = (long long)N * HW;
  float lo = INFINITY, hi = 0.0f;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long n = idx / HW, p = idx - n * HW;
    const int y = (int)(p / W), x = (int)(p - (long long)y * W);
    const float* l = logits + n * 3 * HW;
    // replicate padding of the difference maps: the last row uses rows (H-2, H-1), the last column columns (W-2, W-1)
    const int yy = H > 1 ? min(y, H - 2) : 0, xx = W > 1 ? min(x, W - 2) : 0;
    float e = 0.0f;
    float a0[3], a1[3], b0[3], b1[3];
    softmax3(l, HW, (long long)yy * W + x, a0); softmax3(l, HW, (long long)min(yy + 1, H - 1) * W + x, a1);
    softmax3(l, HW, (long long)y * W + xx, b0); softmax3(l, HW, (long long)y * W + min(xx + 1, W - 1), b1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float dy = fabsf(a1[c] - a0[c]), dx = fabsf(b1[c] - b0[c]);
      e += sqrtf(dy * dy + dx * dx);
    }
    e /= 3.0f;
    edges[idx] = e;
    lo = fminf(lo, e); hi = fmaxf(hi, e);
  }
  for (int o = 16; o; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
  if ((threadIdx.x & 31) == 0 && lo <= hi) { atomicMin(minmax, __float_as_uint(lo)); atomicMax(minmax + 1, __float_as_uint(hi)); }
}

// refined = logits + blend * edge_conv(logits) * normalised_edges   (:131-149)
__global__ void boundary_blend_kernel(const float* __restrict__ logits, const float* __restrict__ corr, const float* __restrict__ edges,
                                      const unsigned int* __restrict__ minmax, const float* __restrict__ blend, int N, long long HW, float* __restrict__ out) {
  const float lo = __uint_as_float(minmax[0]), hi = __uint_as_float(minmax[1]);
  const bool flat = (hi - lo) < 1e-6f;
  const float inv = 1.0f / ((hi - lo) + 1e-6f), bw = *blend;
  const long long total = (long long)N * 3 * HW;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long n = idx / (3 * HW), p = idx % HW;
    const float e = flat ? 0.0f : (edges[n * HW + p] - lo) * inv;
    out[idx] = logits[idx] + bw * corr[idx] * e;
  }
}

// NHWC fp16 slice -> NCHW fp32 (aux outputs: shared_features, fg_attention)
__global__ void nhwc_half_to_nchw_float_kernel(const __half* __restrict__ in, int N, int HW, int C, int cs, int lo, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (p < HW && c < C) ? his_ld1(in + ((long long)n * HW + p) * cs + c, lo) : 0.0f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (p < HW && c < C) out[((long long)n * C + c) * HW + p] = tile[threadIdx.x][i];
  }
}

// ------------------------------------------------------------------------------------ UNet input / output
// PreTrainedPeopleSegmentationUNet.normalize_input (..._unet.py:1885-1890): x/255 iff x.max() > 1, then (x-mean)/std.
// The max is reduced on the device (no host sync); affine = [a0,a1,a2,b0,b1,b2] with x*a+b.
__global__ void max_reduce_kernel(const float* __restrict__ x, long long n, unsigned int* __restrict__ flag) {
  unsigned int over = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) over |= (x[i] > 1.0f);
  over = __any_sync(0xffffffffu, over);
  if ((threadIdx.x & 31) == 0 && over) atomicOr(flag, 1u);
}
__global__ void input_affine_kernel(const unsigned int* __restrict__ flag, float m0, float m1, float m2, float s0, float s1, float s2,
                                    float* __restrict__ affine) {
  const float d = *flag ? 255.0f : 1.0f;
  const float m[3] = {m0, m1, m2}, s[3] = {s0, s1, s2};
  for (int c = 0; c < 3; ++c) { affine[c] = 1.0f / (d * s[c]); affine[3 + c] = -m[c] / s[c]; }
}

// Stem input for the tensor cores: normalise (x*a[c]+b[c]) and space-to-depth the NCHW fp32 image into NHWC fp16
// [N, H/2, W/2, 16] with channel (sy*2+sx)*3 + c (12 used, 4 zero).  A 3x3 stride-2 pad-1 conv over the image is then a 2x2
// stride-1 conv over this tensor (taps (-1,-1),(-1,0),(0,-1),(0,0)), which the halo-mode GEMM runs as a 3x3 with five zero taps.
__global__ void s2d_input_kernel(const float* __restrict__ img, int N, int H, int W, const float* __restrict__ affine, __half* __restrict__ out, int split) {
  const int Ho = H >> 1, Wo = W >> 1;
  const long long total = (long long)N * Ho * Wo;
  const float a0 = affine[0], a1 = affine[1], a2 = affine[2], b0 = affine[3], b1 = affine[4], b2 = affine[5];
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % Wo), oy = (int)((idx / Wo) % Ho);
    const long long n = idx / ((long long)Wo * Ho);
    const float* p0 = img + (n * 3) * (long long)H * W + (long long)(2 * oy) * W + 2 * ox;
    float v[16];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const float* q = p0 + (s >> 1) * W + (s & 1);
      v[s * 3 + 0] = fmaf(__ldg(q), a0, b0);
      v[s * 3 + 1] = fmaf(__ldg(q + (long long)H * W), a1, b1);
      v[s * 3 + 2] = fmaf(__ldg(q + 2LL * H * W), a2, b2);
    }
#pragma unroll
    for (int e = 12; e < 16; ++e) v[e] = 0.0f;
    // split: pixel = [16 hi | 16 lo]
    __half* dst = out + idx * (split ? 32 : 16);
    his_st8(dst, split ? 16 : 0, v);
    his_st8(dst + 8, split ? 16 : 0, v + 8);
  }
}

// output_conv 1x1 (1->2) + export-wrapper binary mask: two = [w0*x+b0, w1*x+b1]; binary = softmax(two)[:,0]
__global__ void unet_outputs_kernel(const float* __restrict__ one, int B, long long HW, float w0, float w1, float b0, float b1,
                                    float* __restrict__ two, float* __restrict__ binary) {
  const long long total = (long long)B * HW;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long n = idx / HW, p = idx % HW;
    const float x = one[idx];
    const float a = w0 * x + b0, b = w1 * x + b1;
    if (two) { two[(n * 2) * HW + p] = a; two[(n * 2 + 1) * HW + p] = b; }
    if (binary) { const float m = fmaxf(a, b); const float ea = expf(a - m), eb = expf(b - m); binary[idx] = ea / (ea + eb); }
  }
}

__global__ void fill_u32_kernel(unsigned int* p, long long n, unsigned int v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace

int his_set_error(int code, const char* msg) {
  snprintf(g_err, sizeof g_err, "%s", msg ? msg : "");
  return code;
}

#define ST ((cudaStream_t)stream)


extern "C" {

const char* his_last_error(void) { return g_err; }

int his_version(void) { return 200; }

int his_roi_align(const void* feat, int feat_is_half, long long sN, long long sC, long long sH, long long sW, int B, int C, int H, int W,
                  const float* rois, int n_rois, int oh, int ow, float scale_h, float scale_w, int aligned, void* out_half, int out_cs,
                  float* out_f32, int split, void* stream) {
  if (n_rois == 0) return HIS_OK;
  if (!feat || (!out_half && !out_f32) || !rois) return his_set_error(HIS_ERR_INVALID_ARG, "roi_align: null pointer");
  if (oh <= 0 || ow <= 0 || C <= 0) return his_set_error(HIS_ERR_INVALID_ARG, "roi_align: bad shape");
  const long long total = (long long)n_rois * oh * ow;
  if (feat_is_half)
    roi_align_kernel<__half><<<grid_for(total), kThreads, 0, ST>>>((const __half*)feat, sN, sC, sH, sW, B, C, H, W, rois, n_rois, oh, ow,
                                                                  scale_h, scale_w, aligned, (__half*)out_half, out_cs, split ? out_cs / 2 : 0, out_f32);
  else
    roi_align_kernel<float><<<grid_for(total), kThreads, 0, ST>>>((const float*)feat, sN, sC, sH, sW, B, C, H, W, rois, n_rois, oh, ow,
                                                                 scale_h, scale_w, aligned, (__half*)out_half, out_cs, split ? out_cs / 2 : 0, out_f32);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_roi_align_fused(const float* feat0, int C0, float scale_h0, float scale_w0, int aligned0, void* out_half0, int out_cs0, float* out_f0,
                        const float* feat1, int C1, float scale_h1, float scale_w1, int aligned1, void* out_half1, int out_cs1, float* out_f1,
                        int B, int H, int W, const float* rois, int n_rois, int oh, int ow, int split, void* stream) {
  if (n_rois == 0) return HIS_OK;
  if (!feat0 || !rois || C0 <= 0 || C1 < 0 || C0 > 3 || C1 > 3 || C0 + C1 > kRoiMaxC || (C1 > 0 && !feat1))
    return his_set_error(HIS_ERR_INVALID_ARG, "roi_align_fused: bad arguments (at most 3 channels per source, 5 over the two)");
  if (oh <= 0 || ow <= 0) return his_set_error(HIS_ERR_INVALID_ARG, "roi_align_fused: bad shape");
  RoiSrc s0{feat0, C0, scale_h0, scale_w0, aligned0, (__half*)out_half0, out_cs0, split ? out_cs0 / 2 : 0, out_f0};
  RoiSrc s1{feat1, C1, scale_h1, scale_w1, aligned1, (__half*)out_half1, out_cs1, split ? out_cs1 / 2 : 0, out_f1};
  const long long rows = (long long)n_rois * oh;
  roi_align_rows_kernel<<<(unsigned)((rows + kRoiWarps - 1) / kRoiWarps), kRoiWarps * 32, 0, ST>>>(s0, s1, B, H, W, rois, n_rois, oh, ow);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_conv_direct(const void* in, int in_fmt, const float* in_affine, int N, int H, int W, int cin, int in_cs, const void* w,
                    const float* scale, const float* shift, int cout, int kh, int kw, int stride, int pad, int act, float act_beta,
                    int res_mode, const void* res, int res_cs, void* out_half, int out_cs, float* out_f32, int split, void* stream) {
  if (!in || !w || !scale || !shift || (!out_half && !out_f32)) return his_set_error(HIS_ERR_INVALID_ARG, "conv_direct: null pointer");
  if (res_mode && !res) return his_set_error(HIS_ERR_INVALID_ARG, "conv_direct: res_mode without residual");
  DirectConvParams p;
  p.in = in; p.in_fmt = in_fmt; p.in_affine = in_affine; p.N = N; p.H = H; p.W = W; p.Cin = cin; p.in_cs = in_cs;
  p.w = w; p.w_f32 = split ? 1 : 0; p.in_lo = (split && in_fmt != 1) ? in_cs / 2 : 0; p.res_lo = split ? res_cs / 2 : 0;
  p.out_lo = split ? out_cs / 2 : 0; p.scale = scale; p.shift = shift; p.Cout = cout; p.kh = kh; p.kw = kw; p.stride = stride; p.pad = pad;
  p.Ho = (H + 2 * pad - kh) / stride + 1; p.Wo = (W + 2 * pad - kw) / stride + 1;
  p.act = act; p.act_beta = act_beta; p.res_mode = res_mode; p.res = (const __half*)res; p.res_cs = res_cs;
  p.out_h = (__half*)out_half; p.out_cs = out_cs; p.out_f = out_f32;
  if (N == 0) return HIS_OK;
  // specialised kernels for the shapes the path actually uses
  if (cin <= 4 && cout % 32 == 0 && out_half && !out_f32 && !res_mode && (out_cs % 8) == 0 && kh == kw && (kh == 1 || kh == 3) &&
      ((size_t)kh * kw * cin + 2) * cout * sizeof(float) <= 48 * 1024 && (in_fmt == 1 || in_fmt == 0)) {
    const long long total = (long long)N * p.Ho * p.Wo * (cout / 32);
    const int g = grid_for(total, 128);
    const size_t sm = ((size_t)kh * kw * cin * cout + 2 * (size_t)cout) * sizeof(float);
    bool done = true;
    if (kh == 3 && cin == 3) small_cin_conv_kernel<9, 3><<<g, 128, sm, ST>>>(p);
    else if (kh == 1 && cin == 2) small_cin_conv_kernel<1, 2><<<g, 128, sm, ST>>>(p);
    else if (kh == 1 && cin == 3) small_cin_conv_kernel<1, 3><<<g, 128, sm, ST>>>(p);
    else done = false;
    if (done) { HIS_CHECK_LAUNCH(); return HIS_OK; }
  }
  if (in_fmt == 2 && !(cout == 1 && kh == 3 && kw == 3 && stride == 1 && pad == 1 && cin == 16 && out_f32 && !out_half && !res_mode &&
                       (H % 2) == 0 && (W % 2) == 0 && in_cs % 8 == 0))
    return his_set_error(HIS_ERR_UNSUPPORTED, "conv_direct: the phase-packed input format exists for the 3x3 16->1 head on even image sizes");
  if ((in_fmt == 0 || in_fmt == 2) && cout <= 3 && out_f32 && !out_half && !res_mode && cin % 8 == 0 && in_cs % 8 == 0 &&
      (size_t)kh * kw * cin * cout * sizeof(float) <= 48 * 1024) {
    const long long total = (long long)N * p.Ho * p.Wo;
    const size_t sm = (size_t)kh * kw * cin * cout * sizeof(float);
    if (cout == 1 && kh == 3 && kw == 3 && stride == 1 && pad == 1 && cin == 16) {
      const long long warps = (long long)N * ((H + kHeadStripRows - 1) / kHeadStripRows) * ((W + kHeadStripCols - 1) / kHeadStripCols);
      head3x3_c1_kernel<16><<<grid_for(warps * 32), kThreads, 0, ST>>>(p);
      HIS_CHECK_LAUNCH();
      return HIS_OK;
    }
    if (cout == 1) small_cout_conv_kernel<1><<<grid_for(total), kThreads, sm, ST>>>(p);
    else if (cout == 3) small_cout_conv_kernel<3><<<grid_for(total), kThreads, sm, ST>>>(p);
    else small_cout_conv_kernel<2><<<grid_for(total), kThreads, sm, ST>>>(p);
    HIS_CHECK_LAUNCH();
    return HIS_OK;
  }
  const int cot = (cout % 8 == 0) ? 8 : (cout % 4 == 0) ? 4 : (cout % 2 == 0) ? 2 : 1;
  const long long total = (long long)N * p.Ho * p.Wo * (cout / cot);
  const int g = grid_for(total);
  switch (cot) {
    case 8: direct_conv_kernel<8><<<g, kThreads, 0, ST>>>(p); break;
    case 4: direct_conv_kernel<4><<<g, kThreads, 0, ST>>>(p); break;
    case 2: direct_conv_kernel<2><<<g, kThreads, 0, ST>>>(p); break;
    default: direct_conv_kernel<1><<<g, kThreads, 0, ST>>>(p); break;
  }
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

// Blocks per image of the pixel reductions.  `fixed_cap` > 0: the count depends on the per-image size only, never on the batch --
// the partial sums of an image (and with them every bit of its result) are then the same whatever batch the image arrives in
// (capacity-bucketed launch plans, sharded batches).  fixed_cap == 0: the register-resident depthwise fallback, sized to the batch.
static int pool_grid_x(long long per_img, int threads, int N, int per_sm, int cgs, int fixed_cap = 0) {
  long long gx = (per_img + threads - 1) / threads;
  const long long cap = fixed_cap > 0 ? fixed_cap : (148LL * per_sm + N - 1) / (N > 0 ? N : 1);
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  const int m = grid_multiple(cgs, threads);
  return (int)((gx + m - 1) / m * m);
}

constexpr int kDwRows = 4;                                   // output rows per thread (RY)
static inline int dw_vec(int k) { return k == 5 ? 4 : 8; }   // channels per thread
static int dw_grid_x(int N, int Ho, int Wo, int C, int k) {
  const int cgs = C / dw_vec(k);
  const long long per_img = (long long)((Ho + kDwRows - 1) / kDwRows) * Wo * cgs;
  return pool_grid_x(per_img, kDwThreads, N, 12, cgs);        // ~12 blocks (4 waves of 3 resident) per SM over the whole batch
}

// tiled kernel geometry: 4 output columns per thread (32-wide tiles) for stride 1 on wide images, 2 otherwise
static inline int dw2_xpt(int Wo, int stride) { return (stride == 1 && Wo >= 64) ? 4 : 2; }
static inline bool dw_use_tiled() { const char* e = getenv("HIS_DW_TILED"); return !e || atoi(e) != 0; }

int his_depthwise_pool_parts(int N, int H, int W, int C, int k, int stride) {
  const int pad = ((stride - 1) + (k - 1)) / 2;
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  if (C <= 0 || C % 8) return 0;
  if (dw_use_tiled()) {
    const int tow = kDw2Strips * dw2_xpt(Wo, stride);
    return ((Wo + tow - 1) / tow) * ((Ho + kDw2Rows - 1) / kDw2Rows);
  }
  return dw_grid_x(N, Ho, Wo, C, k);
}

int his_pool_sum_parts(int N, int HW, int C) {
  const int threads = threads_multiple_of(C / 8);
  if (threads == 0) return 0;
  return pool_grid_x(((long long)HW * (C / 8) + 31) / 32, threads, N, 8, C / 8, 64);
}

int his_depthwise_conv(const void* in, int N, int H, int W, int C, int in_cs, const void* w, const float* scale, const float* shift, int k,
                       int stride, int act, void* out, int out_cs, float* pool_sums, int split, void* stream) {
  if (!in || !w || !scale || !shift || !out) return his_set_error(HIS_ERR_INVALID_ARG, "depthwise: null pointer");
  if (C % 8 || in_cs % 8 || out_cs % 8) return his_set_error(HIS_ERR_UNSUPPORTED, "depthwise: channels must be multiples of 8");
  if (N == 0) return HIS_OK;
  if (!((k == 3 || k == 5) && (stride == 1 || stride == 2)))
    return his_set_error(HIS_ERR_UNSUPPORTED, "depthwise: kernel size must be 3 or 5, stride 1 or 2");
  if (act != HIS_ACT_SILU && act != HIS_ACT_NONE) return his_set_error(HIS_ERR_UNSUPPORTED, "depthwise: activation must be SiLU or none");
  DwParams p;
  p.in = (const __half*)in; p.N = N; p.H = H; p.W = W; p.C = C; p.in_cs = in_cs; p.w = (const __half*)w; p.scale = scale; p.shift = shift;
  p.k = k; p.stride = stride; p.pad = ((stride - 1) + (k - 1)) / 2; p.Ho = (H + 2 * p.pad - k) / stride + 1; p.Wo = (W + 2 * p.pad - k) / stride + 1;
  p.act = act; p.out = (__half*)out; p.out_cs = out_cs; p.pool = pool_sums;
  if ((long long)p.Ho * p.Wo * (C / 4) >= (1LL << 31)) return his_set_error(HIS_ERR_UNSUPPORTED, "depthwise: image too large");
  if (split) {      // split-fp16 activations, fp32 taps (w = [k*k][C] float): the tiled kernel's geometry, operands from global memory
    if (!dw_use_tiled() || N > 65535 || (in_cs % 16) || (out_cs % 16)) return his_set_error(HIS_ERR_UNSUPPORTED, "depthwise (split): unsupported configuration");
    const int xpt = dw2_xpt(p.Wo, stride), tow = kDw2Strips * xpt;
    const int tiles_x = (p.Wo + tow - 1) / tow, tiles_y = (p.Ho + kDw2Rows - 1) / kDw2Rows;
    dim3 g2(tiles_x * tiles_y, (C + kDw2Cb - 1) / kDw2Cb, N);
    const float* w32 = (const float*)w;
    const int il = in_cs / 2, ol = out_cs / 2;
#define DWS_LAUNCH(K_, S_, X_)                                                                                            \
  do {                                                                                                                    \
    constexpr int smem = DwsCfg<K_, S_, X_>::kSmem;                                                                       \
    if (act == HIS_ACT_SILU) {                                                                                            \
      static PerDeviceOnce attr_a;                                                                                        \
      if (attr_a.first()) cudaFuncSetAttribute(depthwise_split_kernel<K_, S_, X_, HIS_ACT_SILU>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
      depthwise_split_kernel<K_, S_, X_, HIS_ACT_SILU><<<g2, kDw2Threads, smem, ST>>>(p, tiles_x, w32, il, ol);           \
    } else {                                                                                                              \
      static PerDeviceOnce attr_b;                                                                                        \
      if (attr_b.first()) cudaFuncSetAttribute(depthwise_split_kernel<K_, S_, X_, HIS_ACT_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
      depthwise_split_kernel<K_, S_, X_, HIS_ACT_NONE><<<g2, kDw2Threads, smem, ST>>>(p, tiles_x, w32, il, ol);           \
    }                                                                                                                     \
  } while (0)
    if (k == 3 && stride == 1 && xpt == 4) DWS_LAUNCH(3, 1, 4);
    else if (k == 3 && stride == 1) DWS_LAUNCH(3, 1, 2);
    else if (k == 5 && stride == 1 && xpt == 4) DWS_LAUNCH(5, 1, 4);
    else if (k == 5 && stride == 1) DWS_LAUNCH(5, 1, 2);
    else if (k == 3) DWS_LAUNCH(3, 2, 2);
    else DWS_LAUNCH(5, 2, 2);
#undef DWS_LAUNCH
    HIS_CHECK_LAUNCH();
    return HIS_OK;
  }
  if (dw_use_tiled() && N <= 65535) {
    const int xpt = dw2_xpt(p.Wo, stride), tow = kDw2Strips * xpt;
    const int tiles_x = (p.Wo + tow - 1) / tow, tiles_y = (p.Ho + kDw2Rows - 1) / kDw2Rows;
    dim3 g2(tiles_x * tiles_y, (C + kDw2Cb - 1) / kDw2Cb, N);
#define DW2_LAUNCH(K_, S_, X_)                                                                                                   \
  do {                                                                                                                           \
    constexpr int smem = Dw2Cfg<K_, S_, X_>::kSmem;                                                                              \
    if (act == HIS_ACT_SILU) {                                                                                                   \
      static PerDeviceOnce attr_a;                                                                                               \
      if (attr_a.first()) cudaFuncSetAttribute(depthwise_tiled_kernel<K_, S_, X_, HIS_ACT_SILU>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
      depthwise_tiled_kernel<K_, S_, X_, HIS_ACT_SILU><<<g2, kDw2Threads, smem, ST>>>(p, tiles_x);                               \
    } else {                                                                                                                     \
      static PerDeviceOnce attr_b;                                                                                               \
      if (attr_b.first()) cudaFuncSetAttribute(depthwise_tiled_kernel<K_, S_, X_, HIS_ACT_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
      depthwise_tiled_kernel<K_, S_, X_, HIS_ACT_NONE><<<g2, kDw2Threads, smem, ST>>>(p, tiles_x);                               \
    }                                                                                                                            \
  } while (0)
    if (k == 3 && stride == 1 && xpt == 4) DW2_LAUNCH(3, 1, 4);
    else if (k == 3 && stride == 1) DW2_LAUNCH(3, 1, 2);
    else if (k == 5 && stride == 1 && xpt == 4) DW2_LAUNCH(5, 1, 4);
    else if (k == 5 && stride == 1) DW2_LAUNCH(5, 1, 2);
    else if (k == 3) DW2_LAUNCH(3, 2, 2);
    else DW2_LAUNCH(5, 2, 2);
#undef DW2_LAUNCH
    HIS_CHECK_LAUNCH();
    return HIS_OK;
  }
  dim3 grid(dw_grid_x(N, p.Ho, p.Wo, C, k), N);
  const size_t sm = pool_sums ? kDwThreads * dw_vec(k) * sizeof(float) : 0;
#define DW_LAUNCH(K_, S_, V_)                                                                                        \
  do {                                                                                                               \
    if (act == HIS_ACT_SILU) depthwise_kernel<K_, S_, HIS_ACT_SILU, V_, kDwRows><<<grid, kDwThreads, sm, ST>>>(p);    \
    else depthwise_kernel<K_, S_, HIS_ACT_NONE, V_, kDwRows><<<grid, kDwThreads, sm, ST>>>(p);                        \
  } while (0)
  if (k == 3 && stride == 1) DW_LAUNCH(3, 1, 8);
  else if (k == 3 && stride == 2) DW_LAUNCH(3, 2, 8);
  else if (k == 5 && stride == 1) DW_LAUNCH(5, 1, 4);
  else DW_LAUNCH(5, 2, 4);
#undef DW_LAUNCH
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_pool_sum(const void* in, int N, int HW, int C, int cs, float* pool_sums, int split, void* stream) {
  if (!in || !pool_sums) return his_set_error(HIS_ERR_INVALID_ARG, "pool_sum: null pointer");
  if (C % 8 || cs % 8) return his_set_error(HIS_ERR_UNSUPPORTED, "pool_sum: channels must be multiples of 8");
  if (N == 0) return HIS_OK;
  const long long per_img = (long long)HW * (C / 8);
  const int threads = threads_multiple_of(C / 8);
  if (threads == 0) return his_set_error(HIS_ERR_UNSUPPORTED, "pool_sum: more than 8192 channels");
  const int gx = pool_grid_x((per_img + 31) / 32, threads, N, 8, C / 8, 64);
  dim3 grid(gx, N);
  pool_sum_kernel<<<grid, threads, threads * 8 * sizeof(float), ST>>>((const __half*)in, HW, C, cs, split ? cs / 2 : 0, pool_sums);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_se_gate(float* pool_sums, int nparts, int N, int HW, int C, int R, const float* w1, const float* b1, const float* w2,
                const float* b2, int act, float act_beta, float* hidden_ws, float* gate, void* stream) {
  if (!pool_sums || !w1 || !w2 || !gate || !hidden_ws) return his_set_error(HIS_ERR_INVALID_ARG, "se_gate: null pointer");
  if (N == 0) return HIS_OK;
  if (nparts < 1) return his_set_error(HIS_ERR_INVALID_ARG, "se_gate: nparts must be >= 1");
  if (nparts > 1) pool_reduce_kernel<<<dim3((C + 31) / 32, N), dim3(32, 32), 0, ST>>>(pool_sums, nparts, C);
  const int wpb = kThreads / 32;
  se_hidden_kernel<<<dim3((R + wpb - 1) / wpb, N), kThreads, 0, ST>>>(pool_sums, nparts, 1.0f / (float)HW, C, R, w1, b1, act, act_beta, hidden_ws);
  se_gate_kernel<<<dim3((C + wpb - 1) / wpb, N), kThreads, 0, ST>>>(hidden_ws, C, R, w2, b2, gate);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_scale_weights(const void* w_packed, const float* gate, int N, long long rows, int K, int C, void* out, int split, void* stream) {
  if (!w_packed || !gate || !out) return his_set_error(HIS_ERR_INVALID_ARG, "scale_weights: null pointer");
  if (K % 8 || C > K) return his_set_error(HIS_ERR_INVALID_ARG, "scale_weights: K must be a multiple of 8 and >= C");
  if (split) {      // rows are [W_hi | W_lo], K/2 elements each
    if (K % 16 || C > K / 2) return his_set_error(HIS_ERR_INVALID_ARG, "scale_weights (split): K must be a multiple of 16 and >= 2*C");
    const long long tot = (long long)N * rows * (K / 16);
    if (tot == 0) return HIS_OK;
    scale_weights_split_kernel<<<grid_for(tot), kThreads, 0, ST>>>((const __half*)w_packed, gate, rows, K / 2, C, tot, (__half*)out);
    HIS_CHECK_LAUNCH();
    return HIS_OK;
  }
  const long long total = (long long)N * rows * (K / 8);
  if (total == 0) return HIS_OK;
  scale_weights_kernel<<<grid_for(total), kThreads, 0, ST>>>((const __half*)w_packed, gate, rows, K, C, total, (__half*)out);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_scale_channels(const void* in, int in_cs, const float* gate, int N, int HW, int C, void* out, int out_cs, int split, void* stream) {
  if (!in || !gate || !out) return his_set_error(HIS_ERR_INVALID_ARG, "scale_channels: null pointer");
  if (C % 8 || in_cs % 8 || out_cs % 8) return his_set_error(HIS_ERR_UNSUPPORTED, "scale_channels: channels must be multiples of 8");
  const long long total = (long long)N * HW * (C / 8);
  if (total == 0) return HIS_OK;
  scale_channels_kernel<<<grid_for(total), kThreads, 0, ST>>>((const __half*)in, in_cs, gate, HW, C, total, (__half*)out, out_cs,
                                                              split ? in_cs / 2 : 0, split ? out_cs / 2 : 0);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_layernorm2d_parts(int N, int HW, int C) {
  (void)N;      // independent of the batch: an image's statistics are summed in the same order whatever batch it arrives in
  long long gx = ((long long)HW * (C / 8) + kThreads * 16 - 1) / (kThreads * 16);
  if (gx > 64) gx = 64;
  return (int)(gx < 1 ? 1 : gx);
}

int his_layernorm2d_act(const void* in, int N, int HW, int C, int in_cs, const float* gamma, const float* beta, float eps, int act,
                        float act_beta, int res_mode, const void* res, int res_cs, double* partials_ws, int nparts_given, void* out, int out_cs,
                        int split, void* stream) {
  if (!in || !gamma || !beta || !partials_ws || !out) return his_set_error(HIS_ERR_INVALID_ARG, "layernorm2d: null pointer");
  if (res_mode && !res) return his_set_error(HIS_ERR_INVALID_ARG, "layernorm2d: res_mode without residual");
  if (C % 8 || in_cs % 8 || out_cs % 8 || (res_mode && res_cs % 8)) return his_set_error(HIS_ERR_UNSUPPORTED, "layernorm2d: channels must be multiples of 8");
  if (N == 0) return HIS_OK;
  // nparts_given > 0: partials_ws already holds that many (sum, sum of squares) pairs per sample, written by the producing GEMM's
  // epilogue (his_conv_gemm_set_ln_partials): the statistics pass over the tensor is skipped
  const int parts = nparts_given > 0 ? nparts_given : his_layernorm2d_parts(N, HW, C);
  const long long per_img_vec = (long long)HW * (C / 8);
  dim3 g1(parts, N);
  if (nparts_given <= 0)
    ln_stats_kernel<<<g1, kThreads, 0, ST>>>((const __half*)in, per_img_vec, HW, C, in_cs, split ? in_cs / 2 : 0, partials_ws);
  long long gx = (per_img_vec + 4 * kThreads - 1) / (4 * kThreads);      // up to four vectors per thread and pass
  const long long cap = (148LL * 16 + N - 1) / N;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  {   // grid stride a multiple of the channel-vector count: every thread of ln_apply_kernel keeps one 8-channel group
    const int cgs = C / 8;
    int a = cgs, b = kThreads;
    while (b) { const int t = a % b; a = b; b = t; }
    const int m = cgs / a;                                  // blocks per period
    gx = (gx + m - 1) / m * m;
  }
  dim3 g2((int)gx, N);
  ln_apply_kernel<<<g2, kThreads, 0, ST>>>((const __half*)in, HW, C, in_cs, partials_ws, parts, gamma, beta, eps, act, act_beta, res_mode,
                                          (const __half*)res, res_cs, (__half*)out, out_cs, split ? in_cs / 2 : 0, split ? res_cs / 2 : 0,
                                          split ? out_cs / 2 : 0);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_groupnorm_parts(int N, int HW, int C) {
  (void)N; (void)C;                                         // independent of the batch (see his_layernorm2d_parts)
  long long parts = (HW + 255) / 256;                       // >= 256 pixels per block
  if (parts > 64) parts = 64;
  return (int)(parts < 1 ? 1 : parts);
}

int his_groupnorm_act(const void* in, int N, int HW, int C, int in_cs, int groups, const float* gamma, const float* beta, float eps, int act,
                      float act_beta, int res_mode, const void* res, int res_cs, float* ws, void* out, int out_cs, int split, void* stream) {
  if (!in || !gamma || !beta || !ws || !out) return his_set_error(HIS_ERR_INVALID_ARG, "groupnorm: null pointer");
  if (res_mode && !res) return his_set_error(HIS_ERR_INVALID_ARG, "groupnorm: res_mode without residual");
  if (C % 8 || in_cs % 8 || out_cs % 8 || (res_mode && res_cs % 8)) return his_set_error(HIS_ERR_UNSUPPORTED, "groupnorm: channels must be multiples of 8");
  if (groups < 1 || C % groups) return his_set_error(HIS_ERR_INVALID_ARG, "groupnorm: channels must be divisible by the group count");
  if (C > kGnMaxC) return his_set_error(HIS_ERR_UNSUPPORTED, "groupnorm: more than 2048 channels");
  if (N == 0) return HIS_OK;
  const int parts = his_groupnorm_parts(N, HW, C);
  const int ppp = (HW + parts - 1) / parts;
  const int cgs = C / 8, PL = kThreads / cgs;
  static PerDeviceOnce attr_done;
  if (attr_done.first()) cudaFuncSetAttribute(gn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const size_t sm1 = (size_t)PL * C * 2 * sizeof(float);      // <= 256/cgs * 8*cgs * 8 B = 16 KB
  gn_stats_kernel<<<dim3(parts, N), kThreads, sm1, ST>>>((const __half*)in, HW, C, in_cs, split ? in_cs / 2 : 0, ppp, parts + 1, ws);
  gn_finalize_kernel<<<dim3((groups + 127) / 128, N), 128, 0, ST>>>(ws, (const __half*)in, in_cs, split ? in_cs / 2 : 0, HW, C, groups, parts, eps);
  const long long per_img_vec = (long long)HW * cgs;
  long long gx = (per_img_vec + kThreads - 1) / kThreads;
  const long long cap = (148LL * 16 + N - 1) / N;
  if (gx > cap) gx = cap;
  gn_apply_kernel<<<dim3((int)(gx < 1 ? 1 : gx), N), kThreads, (size_t)C * 3 * sizeof(float), ST>>>(
      (const __half*)in, HW, C, in_cs, ws, parts, gamma, beta, act, act_beta, res_mode, (const __half*)res, res_cs, (__half*)out, out_cs,
      split ? in_cs / 2 : 0, split ? res_cs / 2 : 0, split ? out_cs / 2 : 0);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_fgaware_norm_act(const void* in, int N, int HW, int C, int in_cs, const float* fg_scale, const float* fg_bias, const float* bg_scale,
                         const float* bg_bias, const float* prob, float eps, int act, float act_beta, int res_mode, const void* res, int res_cs,
                         float* ws, void* out, int out_cs, int split, void* stream) {
  if (!in || !fg_scale || !fg_bias || !bg_scale || !bg_bias || !prob || !ws || !out) return his_set_error(HIS_ERR_INVALID_ARG, "fgaware_norm: null pointer");
  if (res_mode && !res) return his_set_error(HIS_ERR_INVALID_ARG, "fgaware_norm: res_mode without residual");
  if (C % 8 || in_cs % 8 || out_cs % 8 || (res_mode && res_cs % 8)) return his_set_error(HIS_ERR_UNSUPPORTED, "fgaware_norm: channels must be multiples of 8");
  if (C > kGnMaxC) return his_set_error(HIS_ERR_UNSUPPORTED, "fgaware_norm: more than 2048 channels");
  if (N == 0) return HIS_OK;
  const int parts = his_groupnorm_parts(N, HW, C);
  const int ppp = (HW + parts - 1) / parts;
  const int cgs = C / 8, PL = kThreads / cgs;
  static PerDeviceOnce attr_done;
  if (attr_done.first()) cudaFuncSetAttribute(gn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const size_t sm1 = (size_t)PL * C * 2 * sizeof(float);
  gn_stats_kernel<<<dim3(parts, N), kThreads, sm1, ST>>>((const __half*)in, HW, C, in_cs, split ? in_cs / 2 : 0, ppp, parts + 1, ws);
  gn_finalize_kernel<<<dim3((C + 127) / 128, N), 128, 0, ST>>>(ws, (const __half*)in, in_cs, split ? in_cs / 2 : 0, HW, C, C, parts, eps);
  const long long per_img_vec = (long long)HW * cgs;
  long long gx = (per_img_vec + kThreads - 1) / kThreads;
  const long long cap = (148LL * 16 + N - 1) / N;
  if (gx > cap) gx = cap;
  static PerDeviceOnce attr2;
  if (attr2.first()) cudaFuncSetAttribute(fgaware_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  fgaware_apply_kernel<<<dim3((int)(gx < 1 ? 1 : gx), N), kThreads, (size_t)C * 6 * sizeof(float), ST>>>(
      (const __half*)in, HW, C, in_cs, ws, parts, fg_scale, fg_bias, bg_scale, bg_bias, prob, act, act_beta, res_mode, (const __half*)res, res_cs,
      (__half*)out, out_cs, split ? in_cs / 2 : 0, split ? res_cs / 2 : 0, split ? out_cs / 2 : 0);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_convT2x2_small(const float* in, int N, int cin, int h, int w, const float* wt, const float* bias, int cout, void* out, int out_cs,
                       int split, void* stream) {
  if (!in || !wt || !out) return his_set_error(HIS_ERR_INVALID_ARG, "convT2x2_small: null pointer");
  const long long total = (long long)N * 4 * h * w * cout;
  if (total == 0) return HIS_OK;
  convT2x2_small_kernel<<<grid_for(total), kThreads, 0, ST>>>(in, N, cin, h, w, wt, bias, cout, (__half*)out, out_cs, split ? out_cs / 2 : 0);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_maxpool2(const void* in, int N, int H, int W, int C, int in_cs, void* out, int out_cs, int split, void* stream) {
  if (!in || !out) return his_set_error(HIS_ERR_INVALID_ARG, "maxpool2: null pointer");
  if (C % 8 || in_cs % 8 || out_cs % 8) return his_set_error(HIS_ERR_UNSUPPORTED, "maxpool2: channels must be multiples of 8");
  const long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
  if (total == 0) return HIS_OK;
  maxpool2_kernel<<<grid_for(total), kThreads, 0, ST>>>((const __half*)in, N, H, W, C, in_cs, (__half*)out, out_cs, split ? in_cs / 2 : 0,
                                                        split ? out_cs / 2 : 0);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_resize_nearest(const void* in, int N, int H, int W, int C, int in_cs, int Ho, int Wo, void* out, int out_cs, int split, void* stream) {
  if (!in || !out) return his_set_error(HIS_ERR_INVALID_ARG, "resize_nearest: null pointer");
  if (C % 8 || in_cs % 8 || out_cs % 8) return his_set_error(HIS_ERR_UNSUPPORTED, "resize_nearest: channels must be multiples of 8");
  const long long total = (long long)N * Ho * Wo * (C / 8);
  if (total == 0) return HIS_OK;
  resize_nearest_kernel<<<grid_for(total), kThreads, 0, ST>>>((const __half*)in, N, H, W, C, in_cs, Ho, Wo, (__half*)out, out_cs,
                                                              split ? in_cs / 2 : 0, split ? out_cs / 2 : 0);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_resize_bilinear_half(const void* in, int N, int H, int W, int C, int in_cs, int Ho, int Wo, void* out, int out_cs, int split,
                             void* stream) {
  if (!in || !out) return his_set_error(HIS_ERR_INVALID_ARG, "resize_bilinear_half: null pointer");
  if (C % 8 || in_cs % 8 || out_cs % 8) return his_set_error(HIS_ERR_UNSUPPORTED, "resize_bilinear_half: channels must be multiples of 8");
  const long long total = (long long)N * Ho * Wo * (C / 8);
  if (total == 0) return HIS_OK;
  resize_bilinear_half_kernel<<<grid_for(total), kThreads, 0, ST>>>((const __half*)in, N, H, W, C, in_cs, Ho, Wo, (__half*)out, out_cs,
                                                                    split ? in_cs / 2 : 0, split ? out_cs / 2 : 0);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_resize_bilinear_f32(const float* in, int NC, int H, int W, int Ho, int Wo, float* out, void* stream) {
  if (!in || !out) return his_set_error(HIS_ERR_INVALID_ARG, "resize_bilinear: null pointer");
  const long long total = (long long)NC * Ho * Wo;
  if (total == 0) return HIS_OK;
  resize_bilinear_f32_kernel<<<grid_for(total), kThreads, 0, ST>>>(in, NC, H, W, Ho, Wo, out);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_spatial_attention(const void* in, int N, int H, int W, int C, int in_cs, const float* w, int k, float* stats_ws, void* out,
                          int out_cs, int split, void* stream) {
  if (!in || !w || !stats_ws || !out) return his_set_error(HIS_ERR_INVALID_ARG, "spatial_attention: null pointer");
  if (C % 8 || in_cs % 8 || out_cs % 8) return his_set_error(HIS_ERR_UNSUPPORTED, "spatial_attention: channels must be multiples of 8");
  const long long pixels = (long long)N * H * W;
  if (pixels == 0) return HIS_OK;
  channel_stats_kernel<<<grid_for(pixels * 32), kThreads, 0, ST>>>((const __half*)in, pixels, C, in_cs, split ? in_cs / 2 : 0, stats_ws);
  HIS_CHECK_LAUNCH();
  spatial_attention_apply_kernel<<<grid_for(pixels * 32), kThreads, 0, ST>>>((const __half*)in, N, H, W, C, in_cs, stats_ws, w, k,
                                                                            (__half*)out, out_cs, split ? in_cs / 2 : 0,
                                                                            split ? out_cs / 2 : 0);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_spatial_gate(const float* stats, int N, int H, int W, const float* w, int k, float* gate, void* stream) {
  if (!stats || !w || !gate) return his_set_error(HIS_ERR_INVALID_ARG, "spatial_gate: null pointer");
  const long long total = (long long)N * H * W;
  if (total == 0) return HIS_OK;
  spatial_gate_kernel<<<grid_for(total), kThreads, 0, ST>>>(stats, N, H, W, w, k, gate);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_upsample_bgfg(const float* low, int N, int h, int w, const float* wt, const float* scale, const float* shift, const float* w1,
                      const float* b1, int act, float act_beta, float* out, void* stream) {
  if (!low || !wt || !scale || !shift || !w1 || !b1 || !out) return his_set_error(HIS_ERR_INVALID_ARG, "upsample_bgfg: null pointer");
  const long long total = (long long)N * h * w;
  if (total == 0) return HIS_OK;
  upsample_bgfg_kernel<<<grid_for(total), kThreads, 0, ST>>>(low, N, h, w, wt, scale, shift, w1, b1, act, act_beta, out);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_head_combine(const float* bgfg, const float* tn, int N, int H, int W, float* logits, void* stream) {
  if (!bgfg || !tn || !logits) return his_set_error(HIS_ERR_INVALID_ARG, "head_combine: null pointer");
  const long long total = (long long)N * H * W;
  if (total == 0) return HIS_OK;
  head_combine_kernel<<<grid_for(total), kThreads, 0, ST>>>(bgfg, tn, N, (long long)H * W, logits);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_map_f32(const float* in, long long total, int op, const float* param, float* out, void* stream) {
  if (!in || !out || (op == 1 && !param)) return his_set_error(HIS_ERR_INVALID_ARG, "map_f32: null pointer");
  if (total == 0) return HIS_OK;
  map_f32_kernel<<<grid_for(total), kThreads, 0, ST>>>(in, total, op, param, out);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_sigmoid_channel(const float* in, int N, int C, int HW, int c, void* out_half, int out_cs, float* out_f32, int split, void* stream) {
  if (!in || (!out_half && !out_f32)) return his_set_error(HIS_ERR_INVALID_ARG, "sigmoid_channel: null pointer");
  if (c < 0 || c >= C) return his_set_error(HIS_ERR_INVALID_ARG, "sigmoid_channel: channel out of range");
  const long long total = (long long)N * HW;
  if (total == 0) return HIS_OK;
  sigmoid_channel_kernel<<<grid_for(total), kThreads, 0, ST>>>(in, C, HW, c, total, (__half*)out_half, out_cs, split ? out_cs / 2 : 0, out_f32);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_scale_pixels(const void* in, int in_cs, const float* attention, const float* fg_prob, long long pixels, int C, void* out, int out_cs,
                     int split, void* stream) {
  if (!in || !attention || !fg_prob || !out) return his_set_error(HIS_ERR_INVALID_ARG, "scale_pixels: null pointer");
  if (C % 8 || in_cs % 8 || out_cs % 8) return his_set_error(HIS_ERR_UNSUPPORTED, "scale_pixels: channels must be multiples of 8");
  const long long total = pixels * (C / 8);
  if (total == 0) return HIS_OK;
  scale_pixels_kernel<<<grid_for(total), kThreads, 0, ST>>>((const __half*)in, in_cs, attention, fg_prob, C, total, (__half*)out, out_cs,
                                                            split ? in_cs / 2 : 0, split ? out_cs / 2 : 0);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_guided_aux(const float* in, int N, int C, int c, int H, int W, int Ho, int Wo, float* mask_out, float* fg_out, float* bgfg_out,
                   void* stream) {
  if (!in || !mask_out || !fg_out || !bgfg_out) return his_set_error(HIS_ERR_INVALID_ARG, "guided_aux: null pointer");
  if (c < 0 || c >= C) return his_set_error(HIS_ERR_INVALID_ARG, "guided_aux: channel out of range");
  const long long total = (long long)N * Ho * Wo;
  if (total == 0) return HIS_OK;
  guided_aux_kernel<<<grid_for(total), kThreads, 0, ST>>>(in, C, c, H, W, Ho, Wo, total, mask_out, fg_out, bgfg_out);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_depth_to_space2_half(const void* in, int N, int h, int w, int C, int in_cs, void* out, int out_cs, int split, void* stream) {
  if (!in || !out) return his_set_error(HIS_ERR_INVALID_ARG, "depth_to_space2: null pointer");
  if (C % 8 || in_cs % 8 || out_cs % 8 || (split ? in_cs / 2 : in_cs) < 4 * C) return his_set_error(HIS_ERR_UNSUPPORTED, "depth_to_space2: channels must be multiples of 8, in_cs >= 4*C");
  const long long total = (long long)N * 4 * h * w * (C / 8);
  if (total == 0) return HIS_OK;
  depth_to_space2_half_kernel<<<grid_for(total), kThreads, 0, ST>>>((const __half*)in, N, h, w, C, in_cs, (__half*)out, out_cs,
                                                                    split ? in_cs / 2 : 0, split ? out_cs / 2 : 0);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_pixel_shuffle2_f32(const float* in, int N, int C, int in_channels, int h, int w, float* out, void* stream) {
  if (!in || !out) return his_set_error(HIS_ERR_INVALID_ARG, "pixel_shuffle2: null pointer");
  if (in_channels < 4 * C) return his_set_error(HIS_ERR_INVALID_ARG, "pixel_shuffle2: in_channels < 4*C");
  const long long total = (long long)N * C * 4 * h * w;
  if (total == 0) return HIS_OK;
  pixel_shuffle2_kernel<<<grid_for(total), kThreads, 0, ST>>>(in, N, C, in_channels, h, w, out);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_boundary_edges(const float* logits, int N, int H, int W, float* edges, unsigned int* minmax_ws, void* stream) {
  if (!logits || !edges || !minmax_ws) return his_set_error(HIS_ERR_INVALID_ARG, "boundary_edges: null pointer");
  fill_u32_kernel<<<1, 32, 0, ST>>>(minmax_ws, 1, 0x7f800000u);      // +inf
  fill_u32_kernel<<<1, 32, 0, ST>>>(minmax_ws + 1, 1, 0u);
  const long long total = (long long)N * H * W;
  if (total > 0) boundary_edges_kernel<<<grid_for(total), kThreads, 0, ST>>>(logits, N, H, W, edges, minmax_ws);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_boundary_blend(const float* logits, const float* correction, const float* edges, const unsigned int* minmax_ws, const float* blend_weight,
                       int N, int H, int W, float* out, void* stream) {
  if (!logits || !correction || !edges || !minmax_ws || !blend_weight || !out) return his_set_error(HIS_ERR_INVALID_ARG, "boundary_blend: null pointer");
  const long long total = (long long)N * 3 * H * W;
  if (total == 0) return HIS_OK;
  boundary_blend_kernel<<<grid_for(total), kThreads, 0, ST>>>(logits, correction, edges, minmax_ws, blend_weight, N, (long long)H * W, out);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_nhwc_half_to_nchw_float(const void* in, int N, int HW, int C, int cs, float* out, int split, void* stream) {
  if (!in || !out) return his_set_error(HIS_ERR_INVALID_ARG, "layout: null pointer");
  if (N == 0) return HIS_OK;
  dim3 grid((HW + 31) / 32, (C + 31) / 32, N), block(32, 8);
  nhwc_half_to_nchw_float_kernel<<<grid, block, 0, ST>>>((const __half*)in, N, HW, C, cs, split ? cs / 2 : 0, out);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_unet_input_affine(const float* images, long long count, const float* mean3, const float* std3, unsigned int* flag_ws,
                          float* affine6, void* stream) {
  if (!images || !mean3 || !std3 || !flag_ws || !affine6) return his_set_error(HIS_ERR_INVALID_ARG, "input_affine: null pointer");
  fill_u32_kernel<<<1, 32, 0, ST>>>(flag_ws, 1, 0u);
  if (count > 0) max_reduce_kernel<<<grid_for(count), kThreads, 0, ST>>>(images, count, flag_ws);
  input_affine_kernel<<<1, 1, 0, ST>>>(flag_ws, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2], affine6);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_s2d_input(const float* images, int N, int H, int W, const float* affine6, void* out_half, int split, void* stream) {
  if (!images || !affine6 || !out_half) return his_set_error(HIS_ERR_INVALID_ARG, "s2d_input: null pointer");
  if ((H & 1) || (W & 1)) return his_set_error(HIS_ERR_UNSUPPORTED, "s2d_input: H and W must be even");
  const long long total = (long long)N * (H >> 1) * (W >> 1);
  if (total == 0) return HIS_OK;
  s2d_input_kernel<<<grid_for(total), kThreads, 0, ST>>>(images, N, H, W, affine6, (__half*)out_half, split ? 1 : 0);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_unet_outputs(const float* one, int B, int H, int W, float w0, float w1, float b0, float b1, float* two, float* binary, void* stream) {
  if (!one || (!two && !binary)) return his_set_error(HIS_ERR_INVALID_ARG, "unet_outputs: null pointer");
  const long long total = (long long)B * H * W;
  if (total == 0) return HIS_OK;
  unet_outputs_kernel<<<grid_for(total), kThreads, 0, ST>>>(one, B, (long long)H * W, w0, w1, b0, b1, two, binary);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_memset_async(void* ptr, int value, long long bytes, void* stream) {
  if (bytes == 0) return HIS_OK;
  if (!ptr) return his_set_error(HIS_ERR_INVALID_ARG, "memset: null pointer");
  if (cudaMemsetAsync(ptr, value, (size_t)bytes, ST) != cudaSuccess) return his_set_error(HIS_ERR_LAUNCH, "cudaMemsetAsync failed");
  return HIS_OK;
}

}  // extern "C"
