// Shared-memory stencil kernels of the post-processing family (SURVEY §8 a16/a17 + BASELINE config 5).
//
// One CTA = one 32 x 32 output tile of one [H,W] plane: the tile plus its halo is staged in shared memory once
// (coalesced rows, zero / reflect padding resolved while loading), every tap is then a shared-memory read.  Multi-stage
// operators (separable blurs, the two bilateral iterations, the fused clean-up chain) keep their intermediates in a second
// shared plane with a shrinking halo, so a mask is read from HBM once and written once.  Intermediates that the reference
// zero-pads are forced to 0 outside the image, which reproduces F.conv2d(padding=...) stage by stage.
// Float arithmetic keeps the reference's operation order (fmaf accumulation in tap order, no fast-math): the final
// thresholds can sit within a few ulp of a tie.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int TW = 32, TH = 32, kThreads = 256;
constexpr int kMaxR = 8;                                   // largest total halo (fused chain: 1 + 3 + 3 = 7)
constexpr int kPlane = (TW + 2 * kMaxR + 1) * (TH + 2 * kMaxR) + 16;   // floats per shared plane (48 rows x odd pitch 49) + slack for 4-wide strips

enum { PAD_ZERO = 0, PAD_REFLECT = 1 };

struct Tile {
  int x0, y0;      // image coordinates of the tile's first output pixel
  int H, W;
};

// grid = (planes * tiles_x, tiles_y): the plane index rides on gridDim.x (2^31-1 blocks), so one call handles any number of
// planes (BASELINE configs[4] streams 1 M masks; gridDim.z would cap a call at 65 535)
__device__ __forceinline__ int tiles_x_of(int W) { return (W + TW - 1) / TW; }
__device__ __forceinline__ int plane_of(int W) { return blockIdx.x / tiles_x_of(W); }
__device__ __forceinline__ Tile tile_of(int H, int W) {
  Tile t; t.x0 = (blockIdx.x % tiles_x_of(W)) * TW; t.y0 = blockIdx.y * TH; t.H = H; t.W = W; return t;
}

// Stages the (TW + 2R) x (TH + 2R) window around the tile into `s` (row pitch TW + 2R).
template <int PAD, bool CLAMP01>
__device__ __forceinline__ void load_window(const float* __restrict__ img, const Tile& t, int R, float* __restrict__ s, int pitch = 0) {
  const int w = TW + 2 * R, h = TH + 2 * R;
  if (pitch == 0) pitch = w;
  for (int i = threadIdx.x; i < w * h; i += kThreads) {
    const int ly = i / w, lx = i - ly * w;
    int y = t.y0 - R + ly, x = t.x0 - R + lx;
    float v = 0.0f;
    if (PAD == PAD_REFLECT) {
      // F.pad(mode='reflect'): -1 -> 1, H -> H-2; positions right/below the padded image (partial tiles) are never used
      if (y < 0) y = -y; if (y >= t.H) y = 2 * t.H - 2 - y;
      if (x < 0) x = -x; if (x >= t.W) x = 2 * t.W - 2 - x;
      if (y >= 0 && y < t.H && x >= 0 && x < t.W) v = img[(long long)y * t.W + x];
    } else if (y >= 0 && y < t.H && x >= 0 && x < t.W) {
      v = img[(long long)y * t.W + x];
    }
    if (CLAMP01) v = fminf(fmaxf(v, 0.0f), 1.0f);
    s[ly * pitch + lx] = v;
  }
}

__device__ __forceinline__ bool inside(const Tile& t, int y, int x) { return y >= 0 && y < t.H && x >= 0 && x < t.W; }

// ---------------------------------------------------------------------------------------------- edge smoothing family
// BinaryMaskEdgeSmoothing (hed/edge_smoothing.py:10-90) on a window: value at local (ly, lx) of a plane with pitch `w`
__device__ __forceinline__ float edge_smooth_at(const float* __restrict__ s, int w, int ly, int lx, float strength) {
  float lap = 0.0f, g = 0.0f;
  const float gk[9] = {1.f / 16, 2.f / 16, 1.f / 16, 2.f / 16, 4.f / 16, 2.f / 16, 1.f / 16, 2.f / 16, 1.f / 16};
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float v = s[(ly + t / 3 - 1) * w + lx + t % 3 - 1];
    lap = fmaf(v, t == 4 ? 8.0f : -1.0f, lap);
    g = fmaf(v, gk[t], g);
  }
  const float e = fabsf(lap) * strength;
  const float wgt = 1.0f / (1.0f + expf(-e));
  return __fadd_rn(__fmul_rn(s[ly * w + lx], __fsub_rn(1.0f, wgt)), __fmul_rn(g, wgt));
}

__global__ void __launch_bounds__(kThreads) edge_smooth_smem_kernel(const float* __restrict__ in, int H, int W, float thr, float strength,
                                                                    float* __restrict__ out) {
  __shared__ float s[(TW + 2) * (TH + 2)];
  const Tile t = tile_of(H, W);
  const long long plane = (long long)plane_of(W) * H * W;
  load_window<PAD_ZERO, false>(in + plane, t, 1, s);
  __syncthreads();
  for (int i = threadIdx.x; i < TW * TH; i += kThreads) {
    const int ly = i / TW, lx = i - ly * TW, y = t.y0 + ly, x = t.x0 + lx;
    if (y < H && x < W) out[plane + (long long)y * W + x] = edge_smooth_at(s, TW + 2, ly + 1, lx + 1, strength) > thr ? 1.0f : 0.0f;
  }
}

// DirectionalEdgeSmoothing, export_edge_smoothing_onnx.py:63-154: Sobel direction -> cos^2 weights over four directional blurs
__global__ void __launch_bounds__(kThreads) edge_directional_kernel(const float* __restrict__ in, int H, int W, float* __restrict__ out) {
  __shared__ float s[(TW + 4) * (TH + 4)];
  constexpr int w = TW + 4;
  const Tile t = tile_of(H, W);
  const long long plane = (long long)plane_of(W) * H * W;
  load_window<PAD_ZERO, false>(in + plane, t, 2, s);
  __syncthreads();
  for (int i = threadIdx.x; i < TW * TH; i += kThreads) {
    const int ly = i / TW, lx = i - ly * TW, y = t.y0 + ly, x = t.x0 + lx;
    if (y >= H || x >= W) continue;
    const float* c = s + (ly + 2) * w + lx + 2;
    auto at = [&](int dy, int dx) { return c[dy * w + dx]; };
    // cross-correlation in tap order (row-major), zero-weight taps skipped like a dense conv would add 0
    float ex = 0.0f, ey = 0.0f;
    ex = fmaf(at(-1, -1), -1.0f, ex); ex = fmaf(at(-1, 1), 1.0f, ex); ex = fmaf(at(0, -1), -2.0f, ex); ex = fmaf(at(0, 1), 2.0f, ex);
    ex = fmaf(at(1, -1), -1.0f, ex); ex = fmaf(at(1, 1), 1.0f, ex);
    ey = fmaf(at(-1, -1), -1.0f, ey); ey = fmaf(at(-1, 0), -2.0f, ey); ey = fmaf(at(-1, 1), -1.0f, ey); ey = fmaf(at(1, -1), 1.0f, ey);
    ey = fmaf(at(1, 0), 2.0f, ey); ey = fmaf(at(1, 1), 1.0f, ey);
    const float mag = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), 1e-8f));
    const float ang = atan2f(ey, ex);
    float bh = 0.0f, bv = 0.0f, b1 = 0.0f, b2 = 0.0f;
    const float k5[5] = {0.1f, 0.2f, 0.4f, 0.2f, 0.1f};
#pragma unroll
    for (int j = 0; j < 5; ++j) { bh = fmaf(at(0, j - 2), k5[j], bh); bv = fmaf(at(j - 2, 0), k5[j], bv); }
    b1 = fmaf(at(-1, -1), 0.1f, b1); b1 = fmaf(at(0, 0), 0.8f, b1); b1 = fmaf(at(1, 1), 0.1f, b1);
    b2 = fmaf(at(-1, 1), 0.1f, b2); b2 = fmaf(at(0, 0), 0.8f, b2); b2 = fmaf(at(1, -1), 0.1f, b2);
    const float ch = cosf(ang), sh = sinf(ang);
    const float c1 = cosf(__fsub_rn(ang, 0.78539816339744830962f)), c2 = cosf(__fadd_rn(ang, 0.78539816339744830962f));
    float wh = __fmul_rn(ch, ch), wv = __fmul_rn(sh, sh), w1 = __fmul_rn(__fmul_rn(c1, c1), 0.5f), w2 = __fmul_rn(__fmul_rn(c2, c2), 0.5f);
    const float ws = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(wh, wv), w1), w2), 1e-8f);
    wh = __fdiv_rn(wh, ws); wv = __fdiv_rn(wv, ws); w1 = __fdiv_rn(w1, ws); w2 = __fdiv_rn(w2, ws);
    const float blurred = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(bh, wh), __fmul_rn(bv, wv)), __fmul_rn(b1, w1)), __fmul_rn(b2, w2));
    const float em = 1.0f / (1.0f + expf(-__fmul_rn(mag, 3.0f)));
    const float sm = __fadd_rn(__fmul_rn(at(0, 0), __fsub_rn(1.0f, em)), __fmul_rn(blurred, em));
    out[plane + (long long)y * W + x] = sm > 0.5f ? 1.0f : 0.0f;
  }
}

// AdaptiveEdgeSmoothing, export_edge_smoothing_onnx.py:157-218; per-plane runtime parameters
__global__ void __launch_bounds__(kThreads) edge_adaptive_kernel(const float* __restrict__ in, int H, int W, const float* __restrict__ blur_strength,
                                                                 const float* __restrict__ edge_sens, const float* __restrict__ final_thr,
                                                                 float* __restrict__ out) {
  __shared__ float s[(TW + 4) * (TH + 4)];
  constexpr int w = TW + 4;
  const Tile t = tile_of(H, W);
  const int n = plane_of(W);
  const long long plane = (long long)n * H * W;
  load_window<PAD_ZERO, false>(in + plane, t, 2, s);
  __syncthreads();
  const float ethr = __fmul_rn(0.5f, edge_sens[n]), bf = __fdiv_rn(blur_strength[n], 3.0f), fthr = final_thr[n];
  for (int i = threadIdx.x; i < TW * TH; i += kThreads) {
    const int ly = i / TW, lx = i - ly * TW, y = t.y0 + ly, x = t.x0 + lx;
    if (y >= H || x >= W) continue;
    const float* c = s + (ly + 2) * w + lx + 2;
    float lap = 0.0f, avg = 0.0f;
#pragma unroll
    for (int k = 0; k < 9; ++k) lap = fmaf(c[(k / 3 - 1) * w + k % 3 - 1], k == 4 ? 8.0f : -1.0f, lap);
#pragma unroll
    for (int k = 0; k < 25; ++k) avg = fmaf(c[(k / 5 - 2) * w + k % 5 - 2], 1.0f / 25.0f, avg);
    const float m = c[0];
    const float em = fabsf(lap) > ethr ? 1.0f : 0.0f;
    const float sm = __fadd_rn(__fmul_rn(m, __fsub_rn(1.0f, bf)), __fmul_rn(avg, bf));
    const float r = __fadd_rn(__fmul_rn(m, __fsub_rn(1.0f, em)), __fmul_rn(sm, em));
    out[plane + (long long)y * W + x] = r > fthr ? 1.0f : 0.0f;
  }
}

__device__ __forceinline__ float rh(float v, bool fp16) { return fp16 ? __half2float(__float2half_rn(v)) : v; }

// OptimizedEdgeSmoothing, export_edge_smoothing_onnx.py:221-318: separable 5-tap binomial blur (two stages), clamp "sigmoid".
// fp16 != 0 rounds every operator output to half like the exported FP16 graph does.
__global__ void __launch_bounds__(kThreads) edge_optimized_kernel(const float* __restrict__ in, int H, int W, int fp16, float* __restrict__ out) {
  __shared__ float s[(TW + 4) * (TH + 4)];
  __shared__ float hb[TW * (TH + 4)];         // horizontal blur, rows y0-2 .. y0+TH+1
  constexpr int w = TW + 4;
  const Tile t = tile_of(H, W);
  const long long plane = (long long)plane_of(W) * H * W;
  load_window<PAD_ZERO, false>(in + plane, t, 2, s);
  __syncthreads();
  const float g5[5] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f};
  const bool f16 = fp16 != 0;
  for (int i = threadIdx.x; i < TW * (TH + 4); i += kThreads) {
    const int ly = i / TW, lx = i - ly * TW;
    const int y = t.y0 - 2 + ly;
    float v = 0.0f;
    if (y >= 0 && y < H) {           // the vertical pass zero-pads the horizontal result outside the image
#pragma unroll
      for (int j = 0; j < 5; ++j) v = fmaf(s[ly * w + lx + j], g5[j], v);
      v = rh(v, f16);
    }
    hb[i] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TW * TH; i += kThreads) {
    const int ly = i / TW, lx = i - ly * TW, y = t.y0 + ly, x = t.x0 + lx;
    if (y >= H || x >= W) continue;
    const float* c = s + (ly + 2) * w + lx + 2;
    float lap = 0.0f, bl = 0.0f;
#pragma unroll
    for (int k = 0; k < 9; ++k) lap = fmaf(c[(k / 3 - 1) * w + k % 3 - 1], k == 4 ? 8.0f : -1.0f, lap);
#pragma unroll
    for (int j = 0; j < 5; ++j) bl = fmaf(hb[(ly + j) * TW + lx], g5[j], bl);
    lap = rh(lap, f16); bl = rh(bl, f16);
    const float e = rh(__fmul_rn(rh(fabsf(lap), f16), 3.0f), f16);
    float em = rh(__fmul_rn(rh(__fadd_rn(e, 0.5f), f16), 0.5f), f16);
    em = fminf(fmaxf(em, 0.0f), 1.0f);
    const float m = c[0];
    const float a = rh(__fmul_rn(m, rh(__fsub_rn(1.0f, em), f16)), f16), b = rh(__fmul_rn(bl, em), f16);
    out[plane + (long long)y * W + x] = rh(__fadd_rn(a, b), f16) > 0.5f ? 1.0f : 0.0f;
  }
}

// MultiClassEdgeSmoothing front end (hed/edge_smoothing.py:136-143): per-class binary masks of [B,C,H,W] predictions
__global__ void class_masks_kernel(const float* __restrict__ pred, int B, int C, long long HW, int use_argmax, int softmax_first,
                                   float* __restrict__ out) {
  const long long total = (long long)B * HW;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long b = idx / HW, p = idx - b * HW;
    const float* src = pred + b * C * HW + p;
    if (use_argmax) {                  // torch.argmax: first maximum
      int best = 0; float bv = src[0];
      for (int c = 1; c < C; ++c) { const float v = src[c * HW]; if (v > bv) { bv = v; best = c; } }
      for (int c = 0; c < C; ++c) out[(b * C + c) * HW + p] = c == best ? 1.0f : 0.0f;
    } else {
      float mx = -INFINITY, sum = 0.0f;
      if (softmax_first) {
        for (int c = 0; c < C; ++c) mx = fmaxf(mx, src[c * HW]);
        for (int c = 0; c < C; ++c) sum += expf(src[c * HW] - mx);
      }
      for (int c = 0; c < C; ++c) {
        const float v = softmax_first ? expf(src[c * HW] - mx) / sum : src[c * HW];
        out[(b * C + c) * HW + p] = v > 0.5f ? 1.0f : 0.0f;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- bilateral family
// BilateralFilter (hed/bilateral_filter.py:9-113): reflect padding, spatial x range Gaussian weights, normalised
__global__ void __launch_bounds__(kThreads) bilateral_exact_kernel(const float* __restrict__ in, int H, int W, const float* __restrict__ spatial, int k,
                                                                   float inv_2sr2, float* __restrict__ out) {
  __shared__ float s[kPlane];
  __shared__ float sk[81];
  const int R = k / 2, w = TW + 2 * R;
  const Tile t = tile_of(H, W);
  const long long plane = (long long)plane_of(W) * H * W;
  load_window<PAD_REFLECT, false>(in + plane, t, R, s);
  for (int i = threadIdx.x; i < k * k; i += kThreads) sk[i] = spatial[i];
  __syncthreads();
  for (int i = threadIdx.x; i < TW * TH; i += kThreads) {
    const int ly = i / TW, lx = i - ly * TW, y = t.y0 + ly, x = t.x0 + lx;
    if (y >= H || x >= W) continue;
    const float* c = s + (ly + R) * w + lx + R;
    const float cv = c[0];
    float wsum = 0.0f;
    for (int ky = 0; ky < k; ++ky)
      for (int kx = 0; kx < k; ++kx) {
        const float d = c[(ky - R) * w + kx - R] - cv;
        wsum += sk[ky * k + kx] * expf(-(d * d) * inv_2sr2);
      }
    const float den = wsum + 1e-8f;
    float acc = 0.0f;
    for (int ky = 0; ky < k; ++ky)
      for (int kx = 0; kx < k; ++kx) {
        const float v = c[(ky - R) * w + kx - R], d = v - cv;
        acc += v * ((sk[ky * k + kx] * expf(-(d * d) * inv_2sr2)) / den);
      }
    out[plane + (long long)y * W + x] = acc;
  }
}

// one iteration of FastBilateralFilter (hed/bilateral_filter.py:116-216): separable Gaussian of x and x^2 (zero padded stage
// by stage), variance -> exp weight -> blend with the input
__global__ void __launch_bounds__(kThreads) bilateral_fast_iter_kernel(const float* __restrict__ in, int H, int W, const float* __restrict__ k1, int k,
                                                                       float inv_2sr2, float* __restrict__ out) {
  __shared__ float s[kPlane];
  __shared__ float h1[TW * (TH + 2 * kMaxR)], h2[TW * (TH + 2 * kMaxR)];
  __shared__ float sk[16];
  const int R = k / 2, w = TW + 2 * R, hh = TH + 2 * R;
  const Tile t = tile_of(H, W);
  const long long plane = (long long)plane_of(W) * H * W;
  load_window<PAD_ZERO, false>(in + plane, t, R, s);
  if ((int)threadIdx.x < k) sk[threadIdx.x] = k1[threadIdx.x];
  __syncthreads();
  for (int i = threadIdx.x; i < TW * hh; i += kThreads) {
    const int ly = i / TW, lx = i - ly * TW, y = t.y0 - R + ly;
    float a = 0.0f, b = 0.0f;
    if (y >= 0 && y < H)
      for (int j = 0; j < k; ++j) { const float v = s[ly * w + lx + j]; a = fmaf(v, sk[j], a); b = fmaf(__fmul_rn(v, v), sk[j], b); }
    h1[i] = a; h2[i] = b;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TW * TH; i += kThreads) {
    const int ly = i / TW, lx = i - ly * TW, y = t.y0 + ly, x = t.x0 + lx;
    if (y >= H || x >= W) continue;
    float f = 0.0f, q = 0.0f;
    for (int j = 0; j < k; ++j) { f = fmaf(h1[(ly + j) * TW + lx], sk[j], f); q = fmaf(h2[(ly + j) * TW + lx], sk[j], q); }
    const float var = fmaxf(__fsub_rn(q, __fmul_rn(f, f)), 0.0f);
    const float ew = expf(-var * inv_2sr2);
    const float c = s[(ly + R) * w + lx + R];
    out[plane + (long long)y * W + x] = __fadd_rn(__fmul_rn(ew, f), __fmul_rn(__fsub_rn(1.0f, ew), c));
  }
}

// EdgePreservingFilter (guided filter, hed/bilateral_filter.py:219-296), stage 1: box statistics -> a, b
__global__ void __launch_bounds__(kThreads) guided_ab_kernel(const float* __restrict__ x, const float* __restrict__ g, int H, int W, int r, float eps,
                                                             float* __restrict__ a_out, float* __restrict__ b_out) {
  __shared__ float sx[kPlane], sg[kPlane];
  const int w = TW + 2 * r, k = 2 * r + 1;
  const float wk = 1.0f / (float)(k * k);
  const Tile t = tile_of(H, W);
  const long long plane = (long long)plane_of(W) * H * W;
  load_window<PAD_ZERO, false>(x + plane, t, r, sx);
  load_window<PAD_ZERO, false>(g + plane, t, r, sg);
  __syncthreads();
  for (int i = threadIdx.x; i < TW * TH; i += kThreads) {
    const int ly = i / TW, lx = i - ly * TW, y = t.y0 + ly, xx = t.x0 + lx;
    if (y >= H || xx >= W) continue;
    float mx = 0.0f, mg = 0.0f, cxg = 0.0f, cgg = 0.0f;
    for (int ky = 0; ky < k; ++ky)
      for (int kx = 0; kx < k; ++kx) {
        const float vx = sx[(ly + ky) * w + lx + kx], vg = sg[(ly + ky) * w + lx + kx];
        mx = fmaf(vx, wk, mx); mg = fmaf(vg, wk, mg); cxg = fmaf(__fmul_rn(vx, vg), wk, cxg); cgg = fmaf(__fmul_rn(vg, vg), wk, cgg);
      }
    const float cov = __fsub_rn(cxg, __fmul_rn(mx, mg)), var = __fsub_rn(cgg, __fmul_rn(mg, mg));
    const float a = __fdiv_rn(cov, __fadd_rn(var, eps));
    a_out[plane + (long long)y * W + xx] = a;
    b_out[plane + (long long)y * W + xx] = __fsub_rn(mx, __fmul_rn(a, mg));
  }
}

// stage 2: out = box(a) * guide + box(b)
__global__ void __launch_bounds__(kThreads) guided_out_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ g, int H,
                                                              int W, int r, float* __restrict__ out) {
  __shared__ float sa[kPlane], sb[kPlane];
  const int w = TW + 2 * r, k = 2 * r + 1;
  const float wk = 1.0f / (float)(k * k);
  const Tile t = tile_of(H, W);
  const long long plane = (long long)plane_of(W) * H * W;
  load_window<PAD_ZERO, false>(a + plane, t, r, sa);
  load_window<PAD_ZERO, false>(b + plane, t, r, sb);
  __syncthreads();
  for (int i = threadIdx.x; i < TW * TH; i += kThreads) {
    const int ly = i / TW, lx = i - ly * TW, y = t.y0 + ly, x = t.x0 + lx;
    if (y >= H || x >= W) continue;
    float ma = 0.0f, mb = 0.0f;
    for (int ky = 0; ky < k; ++ky)
      for (int kx = 0; kx < k; ++kx) { ma = fmaf(sa[(ly + ky) * w + lx + kx], wk, ma); mb = fmaf(sb[(ly + ky) * w + lx + kx], wk, mb); }
    const long long o = plane + (long long)y * W + x;
    out[o] = __fadd_rn(__fmul_rn(ma, g[o]), mb);
  }
}

// ---------------------------------------------------------------------------------------------- fused clean-up chain
// BASELINE config 5 on full-image masks: BinaryMaskEdgeSmoothing -> BinaryMaskBilateralFilter (n iterations, k x k) in ONE
// pass: the mask is read once and written once.  Every stage is computed on the region the next stage needs (halo shrinking
// from 1 + n*R to 0) and zeroed outside the image, which is exactly the stage-by-stage zero padding of the reference chain.
__device__ __forceinline__ float bilateral_at(const float* __restrict__ s, int w, int ly, int lx, const float* __restrict__ gk, int k) {
  const int R = k / 2;
  float f = 0.0f, f2 = 0.0f;
  for (int ky = 0; ky < k; ++ky)
    for (int kx = 0; kx < k; ++kx) {
      const float v = s[(ly + ky - R) * w + lx + kx - R], wgt = gk[ky * k + kx];
      f = fmaf(v, wgt, f);
      f2 = fmaf(__fmul_rn(v, v), wgt, f2);
    }
  const float c = s[ly * w + lx];
  const float var = fmaxf(__fsub_rn(f2, __fmul_rn(f, f)), 0.0f);
  const float ew = expf(__fmul_rn(-var, 10.0f));
  return __fadd_rn(__fmul_rn(ew, f), __fmul_rn(__fsub_rn(1.0f, ew), c));
}

// Four horizontally adjacent outputs per thread: a window row is loaded once (K + 3 values) and feeds all four accumulators,
// so the kernel leaves the shared-memory-load bound of the one-output form (K*K loads per output -> K*(K+3)/4).  The
// accumulation order per output is unchanged (ky major, kx minor) -> bit-identical to bilateral_at.  BIN: the input is
// binary, v*v == v exactly, so the second moment equals the first.
template <int K, bool BIN>
__device__ __forceinline__ void bilateral4(const float* __restrict__ s, int w, int ly, int lx, const float* __restrict__ gk, float* __restrict__ o) {
  constexpr int R = K / 2;
  float f[4] = {0.0f, 0.0f, 0.0f, 0.0f}, f2[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
  for (int ky = 0; ky < K; ++ky) {
    const float* row = s + (ly + ky - R) * w + lx - R;
    float v[K + 3], v2[K + 3];
#pragma unroll
    for (int i = 0; i < K + 3; ++i) { v[i] = row[i]; if (!BIN) v2[i] = __fmul_rn(v[i], v[i]); }
#pragma unroll
    for (int kx = 0; kx < K; ++kx) {
      const float wgt = gk[ky * K + kx];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f[j] = fmaf(v[kx + j], wgt, f[j]);
        if (!BIN) f2[j] = fmaf(v2[kx + j], wgt, f2[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float c = s[ly * w + lx + j];
    const float q = BIN ? f[j] : f2[j];
    const float var = fmaxf(__fsub_rn(q, __fmul_rn(f[j], f[j])), 0.0f);
    const float ew = expf(__fmul_rn(-var, 10.0f));
    o[j] = __fadd_rn(__fmul_rn(ew, f[j]), __fmul_rn(__fsub_rn(1.0f, ew), c));
  }
}

template <bool BIN>
__device__ __forceinline__ void bilateral4_k(const float* s, int w, int ly, int lx, const float* gk, int k, float* o) {
  switch (k) {
    case 3: bilateral4<3, BIN>(s, w, ly, lx, gk, o); break;
    case 5: bilateral4<5, BIN>(s, w, ly, lx, gk, o); break;
    case 7: bilateral4<7, BIN>(s, w, ly, lx, gk, o); break;
    case 9: bilateral4<9, BIN>(s, w, ly, lx, gk, o); break;
    default:
      for (int j = 0; j < 4; ++j) o[j] = bilateral_at(s, w, ly, lx + j, gk, k);
  }
}

// one bilateral iteration over the region with margin m around the tile (window halo `halo`, pitch w)
template <bool BIN>
__device__ __forceinline__ void bilateral_stage(const float* __restrict__ src, float* __restrict__ dst, const Tile& t, int halo, int m, int w,
                                                const float* __restrict__ gk, int k, bool last, float thr, float* __restrict__ out_plane) {
  const int rw = TW + 2 * m, rh = TH + 2 * m, strips = (rw + 3) / 4;
  for (int i = threadIdx.x; i < strips * rh; i += kThreads) {
    const int ry = i / strips, rx = (i - ry * strips) * 4;
    const int ly = ry + halo - m, lx = rx + halo - m;
    const int y = t.y0 - halo + ly;
    if (y < 0 || y >= t.H) {                       // whole strip outside the image: zero padding of the next stage
      if (!last) for (int j = 0; j < 4 && rx + j < rw; ++j) dst[ly * w + lx + j] = 0.0f;
      continue;
    }
    float o[4];
    bilateral4_k<BIN>(src, w, ly, lx, gk, k, o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (rx + j >= rw) break;
      const int x = t.x0 - halo + lx + j;
      const bool in_img = x >= 0 && x < t.W;
      if (!last) dst[ly * w + lx + j] = in_img ? o[j] : 0.0f;
      else if (in_img) out_plane[(long long)y * t.W + x] = o[j] > thr ? 1.0f : 0.0f;
    }
  }
}

__global__ void __launch_bounds__(kThreads) mask_cleanup_fused_kernel(const float* __restrict__ in, int H, int W, float es_thr, float es_strength,
                                                                      const float* __restrict__ gauss, int k, int iterations, float thr,
                                                                      float* __restrict__ out) {
  __shared__ float pa[kPlane], pb[kPlane];
  __shared__ float gk[81];
  const int R = k / 2;
  const int halo = 1 + iterations * R;                 // <= kMaxR (host checks)
  // odd row pitch: a warp of the bilateral stage covers 4 rows x 8 strips of 4 floats -- with an even pitch rows r and r + 2 fall
  // on the same banks (2-way conflict on every window load: 58.6 M conflict cycles per 64 masks under ncu), an odd pitch spreads
  // the four rows over the four residues mod 4
  const int w = (TW + 2 * halo) | 1;
  const Tile t = tile_of(H, W);
  const long long plane = (long long)plane_of(W) * H * W;
  load_window<PAD_ZERO, false>(in + plane, t, halo, pa, w);
  for (int i = threadIdx.x; i < k * k; i += kThreads) gk[i] = gauss[i];
  __syncthreads();
  // stage 0: edge smoothing on the window shrunk by 1 (the bilateral filter clamps its input to [0,1]: a no-op on {0,1})
  float* src = pa; float* dst = pb;
  int m = halo - 1;                                    // margin still needed around the tile after this stage
  for (int i = threadIdx.x; i < (TW + 2 * m) * (TH + 2 * m); i += kThreads) {
    const int ry = i / (TW + 2 * m), rx = i - ry * (TW + 2 * m);
    const int ly = ry + halo - m, lx = rx + halo - m;
    const int y = t.y0 - halo + ly, x = t.x0 - halo + lx;
    dst[ly * w + lx] = inside(t, y, x) ? (edge_smooth_at(src, w, ly, lx, es_strength) > es_thr ? 1.0f : 0.0f) : 0.0f;
  }
  __syncthreads();
  for (int it = 0; it < iterations; ++it) {
    float* tmp = src; src = dst; dst = tmp;
    m -= R;
    const bool last = it == iterations - 1;
    if (it == 0) bilateral_stage<true>(src, dst, t, halo, m, w, gk, k, last, thr, out + plane);     // input of iteration 0 is {0,1}
    else bilateral_stage<false>(src, dst, t, halo, m, w, gk, k, last, thr, out + plane);
    __syncthreads();
  }
}

// BinaryMaskBilateralFilter alone (hed/bilateral_filter.py:299-406), all iterations in one pass over shared memory
__global__ void __launch_bounds__(kThreads) binary_bilateral_smem_kernel(const float* __restrict__ in, int H, int W, const float* __restrict__ gauss, int k,
                                                                         int iterations, float thr, float* __restrict__ out) {
  __shared__ float pa[kPlane], pb[kPlane];
  __shared__ float gk[81];
  const int R = k / 2;
  const int halo = iterations * R;
  const int w = (TW + 2 * halo) | 1;                   // odd pitch (see mask_cleanup_fused_kernel)
  const Tile t = tile_of(H, W);
  const long long plane = (long long)plane_of(W) * H * W;
  load_window<PAD_ZERO, true>(in + plane, t, halo, pa, w);
  for (int i = threadIdx.x; i < k * k; i += kThreads) gk[i] = gauss[i];
  __syncthreads();
  float* src = pa; float* dst = pb;
  int m = halo;
  for (int it = 0; it < iterations; ++it) {
    m -= R;
    const bool last = it == iterations - 1;
    bilateral_stage<false>(src, dst, t, halo, m, w, gk, k, last, thr, out + plane);
    __syncthreads();
    float* tmp = src; src = dst; dst = tmp;
  }
}

// ---------------------------------------------------------------------------------------------- fused clean-up chain, wide form
// Same chain and the same arithmetic per pixel as mask_cleanup_fused_kernel (which stays as the general form and as the checker
// of this one), laid out for the two limits ncu showed on the first form -- instruction issue (3/4 of the slots were scalar
// shared-memory loads, weight loads and index arithmetic: 28 % FMA-pipe active) and, once those were gone, shared-memory
// wavefronts (every output row re-read its K window rows):
//  * 64 x 32 tile: the stage-1 halo overhead drops from 1.41x to 1.30x;
//  * every stage keeps ITS region at origin (0,0) of its plane, so the strip of 4 outputs at columns [4j, 4j+4) reads input
//    columns [4j, 4j+K+3) -- 16-byte aligned: 3 LDS.128 per window row (K = 7) instead of 10 scalar loads;
//  * a thread owns a 4 x 2 block of outputs: the K+1 window rows it loads feed both output rows (8/14 of the wavefronts), and
//    the two rows share one packed FFMA2 (fma.rn.f32x2, sm_100): (f_row0, f_row1) += v * (w[r][kx], w[r-1][kx]) -- the window
//    value is the broadcast operand, the weight pair comes from a [K+1][K] table of pairs whose out-of-range halves are 0
//    (v * 0 added to a non-negative sum changes nothing, so each output still sees exactly its K*K fmaf in ky-major order);
//  * items are numbered strip-minor: with pitch 84 (two rows = 42 quads = 2 mod 8) any 8 consecutive items of an 18- or
//    16-strip stage touch 8 different bank quads -> conflict-free LDS.128 / STS.128, and the global stores are row segments;
//  * squares of the stage-1 output are written once to a third plane (1 FMUL + store per pixel instead of 17.5 FMUL per output);
//  * edge smoothing of a {0,1} window depends only on (centre, #edge neighbours, #corner neighbours): a 50-entry table per CTA,
//    computed by edge_smooth_at itself on synthetic 3x3 windows (all its sums are exact on such inputs, so the arrangement of
//    the neighbours cannot matter), replaces 18 FMA + sigmoid per pixel.  A window holding anything but 0 / 1 (checked while
//    loading, block-wide vote) takes the general per-pixel path.
// Results are bit-identical to the general form (tests/test_gpu_post.py compares against the unfused per-pixel kernels on
// random, blob and soft masks).
constexpr int TW2 = 64, TH2 = 32, kHalo2 = 7;
constexpr int kWideThreads = 256;   // 352 (stage 1's 18 x 19 blocks in one round, 3 CTAs/SM) measured 6 % slower
constexpr int kPitch2 = 84;                                   // >= 4 * ceil((TW2 + 2*kHalo2 - 2) / 4) + 8, and = 20 mod 32
constexpr int kPlane2 = (TH2 + 2 * kHalo2) * kPitch2;         // 46 rows

struct Region {
  int rw, rh;      // extent of the stage's output region (rh even)
  int oy, ox;      // image coordinates of its (0,0)
};

// (d0, d1) += a * (b0, b1): one FFMA2, `a` as the broadcast operand
__device__ __forceinline__ void ffma2_bcast(float& d0, float& d1, float a, float b0, float b1) {
  unsigned long long ra, rb, rc;
  asm("mov.b64 %0, {%1, %1};" : "=l"(ra) : "f"(a));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b0), "f"(b1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(d0), "f"(d1));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(rc) : "l"(ra), "l"(rb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(rc));
}

__device__ __forceinline__ void load_quads(const float* __restrict__ p, float* __restrict__ v, int n) {
#pragma unroll
  for (int l = 0; l < 3; ++l)
    if (l < n) { const float4 t = reinterpret_cast<const float4*>(p)[l]; v[4 * l] = t.x; v[4 * l + 1] = t.y; v[4 * l + 2] = t.z; v[4 * l + 3] = t.w; }
}

// wp: [K+1][8] pairs (w[r][kx], w[r-1][kx]) as float2, rows outside 0..K-1 are 0
template <int K, bool BIN, bool LAST, typename T>
__device__ __forceinline__ void bilateral_stage2(const float* __restrict__ src, const float* __restrict__ src2, float* __restrict__ dst,
                                                 float* __restrict__ dst2, const float* __restrict__ wp, const Region g, int H, int W, float thr,
                                                 T* __restrict__ out_plane, bool vec_store) {
  constexpr int R = K / 2, NL = (K + 3 + 3) / 4;              // float4 loads per window row
  const int S = (g.rw + 3) >> 2;
  for (int i = threadIdx.x; i < S * (g.rh >> 1); i += kWideThreads) {
    const int rp = i / S, sx = i - rp * S;
    const int c0 = sx * 4, ry = 2 * rp, y = g.oy + ry;
    const bool in0 = y >= 0 && y < H, in1 = y + 1 >= 0 && y + 1 < H;
    if (!in0 && !in1) {                                       // outside the image: zero padding of the next stage
      if (!LAST) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          *reinterpret_cast<float4*>(dst + (ry + o) * kPitch2 + c0) = z;
          *reinterpret_cast<float4*>(dst2 + (ry + o) * kPitch2 + c0) = z;
        }
      }
      continue;
    }
    float f[4][2], f2[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) { f[j][0] = f[j][1] = 0.0f; f2[j][0] = f2[j][1] = 0.0f; }
#pragma unroll 1
    for (int r = 0; r <= K; ++r) {
      float v[12], q[12], w[16];
      load_quads(src + (ry + r) * kPitch2 + c0, v, NL);
      if (!BIN) load_quads(src2 + (ry + r) * kPitch2 + c0, q, NL);
#pragma unroll
      for (int l = 0; l < (2 * K + 3) / 4; ++l) {
        const float4 t = reinterpret_cast<const float4*>(wp + r * 16)[l];
        w[4 * l] = t.x; w[4 * l + 1] = t.y; w[4 * l + 2] = t.z; w[4 * l + 3] = t.w;
      }
#pragma unroll
      for (int kx = 0; kx < K; ++kx)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ffma2_bcast(f[j][0], f[j][1], v[kx + j], w[2 * kx], w[2 * kx + 1]);
          if (!BIN) ffma2_bcast(f2[j][0], f2[j][1], q[kx + j], w[2 * kx], w[2 * kx + 1]);
        }
    }
    const int x = g.ox + c0;
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      float res[4], ctr[4];
      // centre values re-read here (two quads) rather than carried through the window loop in 8 registers
#pragma unroll
      for (int j = 0; j < 4; ++j) ctr[j] = src[(ry + o + R) * kPitch2 + c0 + R + j];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float fj = f[j][o], qq = BIN ? fj : f2[j][o];
        const float var = fmaxf(__fsub_rn(qq, __fmul_rn(fj, fj)), 0.0f);
        const float ew = expf(__fmul_rn(-var, 10.0f));
        res[j] = __fadd_rn(__fmul_rn(ew, fj), __fmul_rn(__fsub_rn(1.0f, ew), ctr[j]));
      }
      const bool yin = o == 0 ? in0 : in1;
      if (!LAST) {
#pragma unroll
        for (int j = 0; j < 4; ++j) res[j] = (yin && x + j >= 0 && x + j < W) ? res[j] : 0.0f;
        *reinterpret_cast<float4*>(dst + (ry + o) * kPitch2 + c0) = make_float4(res[0], res[1], res[2], res[3]);
        *reinterpret_cast<float4*>(dst2 + (ry + o) * kPitch2 + c0) =
            make_float4(__fmul_rn(res[0], res[0]), __fmul_rn(res[1], res[1]), __fmul_rn(res[2], res[2]), __fmul_rn(res[3], res[3]));
      } else if (yin) {
        T* orow = out_plane + (long long)(y + o) * W + x;
#pragma unroll
        for (int j = 0; j < 4; ++j) res[j] = res[j] > thr ? 1.0f : 0.0f;
        if (vec_store && x + 3 < W) {
          if (sizeof(T) == 4) *reinterpret_cast<float4*>(orow) = make_float4(res[0], res[1], res[2], res[3]);
          else *reinterpret_cast<uchar4*>(orow) = make_uchar4((unsigned char)res[0], (unsigned char)res[1], (unsigned char)res[2], (unsigned char)res[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) if (x + j < W) orow[j] = (T)res[j];
        }
      }
    }
  }
}

// 4 CTAs per SM: the window-row loop stays rolled (55 registers; unrolled it wants 128 and spills at 80)
// T = float (the reference's mask tensors) or unsigned char (0 / 1 masks: a quarter of the bytes over PCIe and HBM; the values are
// converted to float while loading, so any other byte value behaves as float(mask) would)
template <int K, int ITS, typename T>
__global__ void __launch_bounds__(kWideThreads, 4) mask_cleanup_wide_kernel(const T* __restrict__ in, int H, int W, float es_thr, float es_strength,
                                                                        const float* __restrict__ gauss, float thr, T* __restrict__ out) {
  __shared__ __align__(16) float pa[kPlane2], pb[kPlane2], pc[kPlane2];
  __shared__ __align__(16) float wp[(K + 1) * 16];
  __shared__ float tab[64];
  constexpr int R = K / 2;
  constexpr int halo = 1 + ITS * R;                           // <= kHalo2 (host checks)
  constexpr int D = 8 - halo;                                 // plane column of the window's first column
  const int tiles_x = (W + TW2 - 1) / TW2;
  const int x0 = (blockIdx.x % tiles_x) * TW2, y0 = blockIdx.y * TH2;
  const long long plane = (long long)(blockIdx.x / tiles_x) * H * W;
  constexpr int W0 = TW2 + 2 * halo, H0 = TH2 + 2 * halo;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool vec_io = (W & 3) == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & (4 * sizeof(T) - 1)) == 0;
  // window, zero padded; plane column p holds image column x0 - 8 + p (80 columns: the window aligned down to 16 bytes, so a
  // row is 20 float4 loads when W % 4 == 0 -- a group of four is then entirely inside or outside the image).  Is every
  // value 0 or 1?
  bool bin = true;
  if (vec_io) {
    constexpr int G = 20;                                     // float4 groups per window row
#pragma unroll
    for (int u = 0; u < (46 * G + kWideThreads - 1) / kWideThreads; ++u) {
      const int i = threadIdx.x + u * kWideThreads;
      const int r = i / G, q = i - r * G;
      const int y = y0 - halo + r, x = x0 - 8 + 4 * q;
      if (r < H0) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (y >= 0 && y < H && x >= 0 && x < W) {
          if (sizeof(T) == 4) v = __ldg(reinterpret_cast<const float4*>(in + plane + (long long)y * W + x));
          else {
            const uchar4 t = __ldg(reinterpret_cast<const uchar4*>(in + plane + (long long)y * W + x));
            v = make_float4((float)t.x, (float)t.y, (float)t.z, (float)t.w);
          }
        }
        bin = bin && (v.x == 0.0f || v.x == 1.0f) && (v.y == 0.0f || v.y == 1.0f) && (v.z == 0.0f || v.z == 1.0f) && (v.w == 0.0f || v.w == 1.0f);
        *reinterpret_cast<float4*>(pa + r * kPitch2 + 4 * q) = v;
      }
    }
    if (threadIdx.x < H0) *reinterpret_cast<float4*>(pa + threadIdx.x * kPitch2 + 80) = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    for (int r = warp; r < H0; r += kWideThreads / 32) {
      const int y = y0 - halo + r;
      const bool yin = y >= 0 && y < H;
      const T* grow = in + plane + (long long)y * W;
      for (int c = lane; c < kPitch2; c += 32) {
        const int x = x0 - 8 + c;
        float v = 0.0f;
        if (yin && c < 80 && x >= 0 && x < W) v = (float)grow[x];
        bin = bin && (v == 0.0f || v == 1.0f);
        pa[r * kPitch2 + c] = v;
      }
    }
  }
  for (int i = threadIdx.x; i < (K + 1) * 16; i += kWideThreads) {                 // pairs (w[r][kx], w[r-1][kx])
    const int r = i >> 4, kx = (i & 15) >> 1, rr = r - (i & 1);
    wp[i] = (kx < K && rr >= 0 && rr < K) ? gauss[rr * K + kx] : 0.0f;
  }
  if (threadIdx.x < 50) {                                     // edge-smoothing table: (centre, #edge, #corner neighbours) -> {0,1}
    const int c = threadIdx.x / 25, ne = (threadIdx.x / 5) % 5, nc = threadIdx.x % 5;
    float* w3 = pc + threadIdx.x * 9;                         // pc is not a plane yet
    w3[4] = (float)c;
    w3[1] = ne > 0 ? 1.0f : 0.0f; w3[3] = ne > 1 ? 1.0f : 0.0f; w3[5] = ne > 2 ? 1.0f : 0.0f; w3[7] = ne > 3 ? 1.0f : 0.0f;
    w3[0] = nc > 0 ? 1.0f : 0.0f; w3[2] = nc > 1 ? 1.0f : 0.0f; w3[6] = nc > 2 ? 1.0f : 0.0f; w3[8] = nc > 3 ? 1.0f : 0.0f;
    tab[threadIdx.x] = edge_smooth_at(w3, 3, 1, 1, es_strength) > es_thr ? 1.0f : 0.0f;
  }
  const int all_bin = __syncthreads_and(bin ? 1 : 0);
  // stage 0: edge smoothing, region shrunk by 1
  {
    constexpr int rw = W0 - 2, rh = H0 - 2;
    const int oy = y0 - halo + 1, ox = x0 - halo + 1;
    if (all_bin) {
      constexpr int S = (rw + 3) >> 2, O = D & 3;
      for (int i = threadIdx.x; i < S * (rh >> 1); i += kWideThreads) {
        const int rp = i / S, sx = i - rp * S, c0 = sx * 4, ry = 2 * rp;
        float w4[4][12];
        const float* base = pa + ry * kPitch2 + c0 + (D & ~3);
#pragma unroll
        for (int r = 0; r < 4; ++r) load_quads(base + r * kPitch2, w4[r], 3);
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          const int y = oy + ry + o;
          const bool yin = y >= 0 && y < H;
          float cs[6], res[4];
#pragma unroll
          for (int l = 0; l < 6; ++l) cs[l] = w4[o][l + O] + w4[o + 2][l + O];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float ne = cs[j + 1] + w4[o + 1][j + O] + w4[o + 1][j + 2 + O], nc = cs[j] + cs[j + 2];
            const int idx = __float2int_rn(fmaf(w4[o + 1][j + 1 + O], 25.0f, fmaf(ne, 5.0f, nc)));
            const int x = ox + c0 + j;
            res[j] = (yin && x >= 0 && x < W) ? tab[idx] : 0.0f;
          }
          *reinterpret_cast<float4*>(pb + (ry + o) * kPitch2 + c0) = make_float4(res[0], res[1], res[2], res[3]);
        }
      }
    } else {
      for (int i = threadIdx.x; i < rw * rh; i += kWideThreads) {
        const int ry = i / rw, c = i - ry * rw, y = oy + ry, x = ox + c;
        const bool in_img = y >= 0 && y < H && x >= 0 && x < W;
        pb[ry * kPitch2 + c] = in_img ? (edge_smooth_at(pa, kPitch2, ry + 1, c + 1 + D, es_strength) > es_thr ? 1.0f : 0.0f) : 0.0f;
      }
    }
  }
  __syncthreads();
  if (ITS == 1) {
    const Region g{TW2, TH2, y0, x0};
    bilateral_stage2<K, true, true, T>(pb, nullptr, nullptr, nullptr, wp, g, H, W, thr, out + plane, vec_io);
  } else {
    const Region g1{TW2 + 2 * R, TH2 + 2 * R, y0 - R, x0 - R};
    bilateral_stage2<K, true, false, T>(pb, nullptr, pa, pc, wp, g1, H, W, thr, nullptr, false);
    __syncthreads();
    const Region g2{TW2, TH2, y0, x0};
    bilateral_stage2<K, false, true, T>(pa, pc, nullptr, nullptr, wp, g2, H, W, thr, out + plane, vec_io);
  }
}

// ---------------------------------------------------------------------------------------------- MaskDilationModule, tiled
// export_hierarchical_instance_peopleseg_onnx.py:85-141: p1 = softmax(logits)[1]; d = max_pool(p1, 2k+1, stride 1, pad k);
// l1 += 2 where d - p1 > 0.1.  p1 is computed ONCE per pixel into shared memory (the per-pixel form evaluated 3 expf + a
// division for each of the (2k+1)^2 neighbours: 1.5 G warp instructions per 5 120 ROIs), the window maximum is taken as a
// row pass then a column pass (max is exact in any order); pixels outside the plane hold -1 and never win, which is
// max_pool2d's implicit -inf padding.  MASK: the instance mask of the dilated logits (instance_mask_kernel's rule) is
// written instead of the logits -- the dilated tensor never goes to HBM.
__device__ __forceinline__ float softmax_p1_of(float l0, float l1, float l2) {
  const float m = fmaxf(l0, fmaxf(l1, l2));
  const float e0 = expf(l0 - m), e1 = expf(l1 - m), e2 = expf(l2 - m);
  return e1 / ((e0 + e1) + e2);
}

constexpr int kDilMax = 16;                                  // largest dilation_pixels (host checks)

// KD = dilation_pixels as a compile-time constant (1 and 2: window extents, divisions and the max loops unroll), 0 = run time
template <bool MASK, int KD>
__global__ void __launch_bounds__(kThreads) dilate_logits_tiled_kernel(const float* __restrict__ logits, int H, int W, int k_rt, float score_thr,
                                                                       float* __restrict__ out, unsigned char* __restrict__ out_u8) {
  __shared__ float sp[(TH + 2 * kDilMax) * (TW + 2 * kDilMax + 1)];      // p1 window, odd pitch
  __shared__ float sr[(TH + 2 * kDilMax) * TW];                          // row maxima
  const int k = KD ? KD : k_rt;
  const Tile t = tile_of(H, W);
  const long long HW = (long long)H * W, n = plane_of(W);
  const float* l0p = logits + n * 3 * HW;
  const int w0 = TW + 2 * k, h0 = TH + 2 * k, pitch = w0 | 1;
  const float inv_w0 = 1.0f / (float)w0;
  // four window pixels per thread and pass: the twelve loads are issued before the first softmax
  for (int base = threadIdx.x; base < w0 * h0; base += 4 * kThreads) {
    float l[4][3];
    int at[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * kThreads;
      const int ly = KD ? i / w0 : (int)(((float)i + 0.5f) * inv_w0), lx = i - ly * w0;   // float form exact for i < 2^13
      const int y = t.y0 - k + ly, x = t.x0 - k + lx;
      at[u] = -1;
      if (i < w0 * h0) {
        at[u] = ly * pitch + lx;
        if (y >= 0 && y < H && x >= 0 && x < W) {
          const float* q = l0p + ((long long)y * W + x);
          l[u][0] = __ldg(q); l[u][1] = __ldg(q + HW); l[u][2] = __ldg(q + 2 * HW);
        } else {
          at[u] = -2 - at[u];                                  // outside the plane: -1 (never the maximum)
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (at[u] >= 0) sp[at[u]] = softmax_p1_of(l[u][0], l[u][1], l[u][2]);
      else if (at[u] < -1) sp[-2 - at[u]] = -1.0f;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TW * h0; i += kThreads) {
    const int ly = i / TW, lx = i % TW;
    const float* row = sp + ly * pitch + lx;
    float m = row[0];
    if (KD) {
#pragma unroll
      for (int d = 1; d <= 2 * KD; ++d) m = fmaxf(m, row[d]);
    } else {
      for (int d = 1; d <= 2 * k; ++d) m = fmaxf(m, row[d]);
    }
    sr[ly * TW + lx] = m;
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < TW * TH / kThreads; ++u) {
    const int i = threadIdx.x + u * kThreads;
    const int ly = i / TW, lx = i % TW;
    const int y = t.y0 + ly, x = t.x0 + lx;
    if (y >= H || x >= W) continue;
    float d = sr[ly * TW + lx];
    if (KD) {
#pragma unroll
      for (int e = 1; e <= 2 * KD; ++e) d = fmaxf(d, sr[(ly + e) * TW + lx]);
    } else {
      for (int e = 1; e <= 2 * k; ++e) d = fmaxf(d, sr[(ly + e) * TW + lx]);
    }
    const float p1 = sp[(ly + k) * pitch + lx + k];
    const float* q = l0p + ((long long)y * W + x);
    const long long p = (long long)y * W + x;
    const float l0 = __ldg(q), l1 = __ldg(q + HW), l2 = __ldg(q + 2 * HW);
    const float l1d = ((d - p1) > 0.1f) ? l1 + 2.0f : l1;
    if (!MASK) {
      float* o = out + n * 3 * HW + p;
      o[0] = l0; o[HW] = l1d; o[2 * HW] = l2;
    } else {
      bool on = (l1d > l0) && (l1d >= l2);
      if (on && score_thr > 0.0f) {
        const float s = expf(l0 - l1d) + 1.0f + expf(l2 - l1d);
        on = (1.0f / s) > score_thr;
      }
      if (out) out[n * HW + p] = on ? 1.0f : 0.0f;
      if (out_u8) out_u8[n * HW + p] = on ? 1 : 0;
    }
  }
}

template <bool MASK>
void launch_dilate(const float* logits, int N, int H, int W, int k, float thr, float* out, unsigned char* out_u8, cudaStream_t st);

inline dim3 tile_grid(int N, int H, int W) { return dim3((unsigned)((long long)N * ((W + TW - 1) / TW)), (H + TH - 1) / TH, 1); }

template <bool MASK>
void launch_dilate(const float* logits, int N, int H, int W, int k, float thr, float* out, unsigned char* out_u8, cudaStream_t st) {
  const dim3 g = tile_grid(N, H, W);
  if (k == 1) dilate_logits_tiled_kernel<MASK, 1><<<g, kThreads, 0, st>>>(logits, H, W, k, thr, out, out_u8);
  else if (k == 2) dilate_logits_tiled_kernel<MASK, 2><<<g, kThreads, 0, st>>>(logits, H, W, k, thr, out, out_u8);
  else dilate_logits_tiled_kernel<MASK, 0><<<g, kThreads, 0, st>>>(logits, H, W, k, thr, out, out_u8);
}

}  // namespace

#define ST ((cudaStream_t)stream)
#define CHECK_PLANES(name)                                                                                              \
  if (N == 0 || H == 0 || W == 0) return HIS_OK;                                                                        \
  if ((long long)N * ((W + TW - 1) / TW) >= (1LL << 31)) return his_set_error(HIS_ERR_UNSUPPORTED, name ": too many planes per call")

namespace {
template <typename T>
bool launch_cleanup_wide(const T* mask, int N, int H, int W, float es_threshold, float es_strength, const float* gauss, int k, int iterations,
                         float threshold, T* out, cudaStream_t st) {
  if (iterations < 1 || iterations > 2 || !(k == 3 || k == 5 || k == 7) || (long long)N * ((W + TW2 - 1) / TW2) >= (1LL << 31)) return false;
  const dim3 grid((unsigned)((long long)N * ((W + TW2 - 1) / TW2)), (H + TH2 - 1) / TH2, 1);
#define HIS_WIDE(K_, I_)                                                                                                            \
  do {                                                                                                                              \
    static PerDeviceOnce carve;                                                                                                     \
    if (carve.first()) cudaFuncSetAttribute(mask_cleanup_wide_kernel<K_, I_, T>, cudaFuncAttributePreferredSharedMemoryCarveout, 100); \
    mask_cleanup_wide_kernel<K_, I_, T><<<grid, kWideThreads, 0, st>>>(mask, H, W, es_threshold, es_strength, gauss, threshold, out); \
  } while (0)
  if (k == 7) { if (iterations == 2) HIS_WIDE(7, 2); else HIS_WIDE(7, 1); }
  else if (k == 5) { if (iterations == 2) HIS_WIDE(5, 2); else HIS_WIDE(5, 1); }
  else { if (iterations == 2) HIS_WIDE(3, 2); else HIS_WIDE(3, 1); }
#undef HIS_WIDE
  return true;
}
}  // namespace

extern "C" {

int his_post_edge_smooth_tiled(const float* mask, int N, int H, int W, float threshold, float blur_strength, float* out, void* stream) {
  if (!mask || !out) return his_set_error(HIS_ERR_INVALID_ARG, "edge_smooth_tiled: null pointer");
  CHECK_PLANES("edge_smooth_tiled");
  edge_smooth_smem_kernel<<<tile_grid(N, H, W), kThreads, 0, ST>>>(mask, H, W, threshold, blur_strength, out);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_post_edge_directional(const float* mask, int N, int H, int W, float* out, void* stream) {
  if (!mask || !out) return his_set_error(HIS_ERR_INVALID_ARG, "edge_directional: null pointer");
  CHECK_PLANES("edge_directional");
  edge_directional_kernel<<<tile_grid(N, H, W), kThreads, 0, ST>>>(mask, H, W, out);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_post_edge_adaptive(const float* mask, int N, int H, int W, const float* blur_strength, const float* edge_sensitivity,
                           const float* final_threshold, float* out, void* stream) {
  if (!mask || !out || !blur_strength || !edge_sensitivity || !final_threshold) return his_set_error(HIS_ERR_INVALID_ARG, "edge_adaptive: null pointer");
  CHECK_PLANES("edge_adaptive");
  edge_adaptive_kernel<<<tile_grid(N, H, W), kThreads, 0, ST>>>(mask, H, W, blur_strength, edge_sensitivity, final_threshold, out);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_post_edge_optimized(const float* mask, int N, int H, int W, int fp16, float* out, void* stream) {
  if (!mask || !out) return his_set_error(HIS_ERR_INVALID_ARG, "edge_optimized: null pointer");
  CHECK_PLANES("edge_optimized");
  edge_optimized_kernel<<<tile_grid(N, H, W), kThreads, 0, ST>>>(mask, H, W, fp16, out);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_post_class_masks(const float* pred, int B, int C, int H, int W, int use_argmax, int softmax_first, float* out, void* stream) {
  if (!pred || !out) return his_set_error(HIS_ERR_INVALID_ARG, "class_masks: null pointer");
  const long long total = (long long)B * H * W;
  if (total == 0 || C == 0) return HIS_OK;
  long long g = (total + 255) / 256; if (g > 148LL * 64) g = 148LL * 64;
  class_masks_kernel<<<(int)g, 256, 0, ST>>>(pred, B, C, (long long)H * W, use_argmax, softmax_first, out);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_post_bilateral_exact(const float* in, int N, int H, int W, const float* spatial_kernel, int k, float sigma_range, float* out, void* stream) {
  if (!in || !out || !spatial_kernel) return his_set_error(HIS_ERR_INVALID_ARG, "bilateral_exact: null pointer");
  if (k < 1 || !(k & 1) || k / 2 > kMaxR) return his_set_error(HIS_ERR_UNSUPPORTED, "bilateral_exact: odd kernel size <= 17");
  if (k > 9) return his_set_error(HIS_ERR_UNSUPPORTED, "bilateral_exact: kernel size <= 9");
  if (H <= k / 2 || W <= k / 2) return his_set_error(HIS_ERR_INVALID_ARG, "bilateral_exact: reflect padding needs H, W > kernel_size/2");
  CHECK_PLANES("bilateral_exact");
  bilateral_exact_kernel<<<tile_grid(N, H, W), kThreads, 0, ST>>>(in, H, W, spatial_kernel, k, 1.0f / (2.0f * sigma_range * sigma_range), out);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_post_bilateral_fast(const float* in, int N, int H, int W, const float* kernel1d, int k, float sigma_range, int iterations, float* ws,
                            float* out, void* stream) {
  if (!in || !out || !kernel1d || (iterations > 1 && !ws)) return his_set_error(HIS_ERR_INVALID_ARG, "bilateral_fast: null pointer");
  if (k < 1 || !(k & 1) || k > 15 || iterations < 1) return his_set_error(HIS_ERR_UNSUPPORTED, "bilateral_fast: odd kernel size <= 15, iterations >= 1");
  CHECK_PLANES("bilateral_fast");
  const float inv = 1.0f / (2.0f * sigma_range * sigma_range);
  const float* src = in;
  for (int it = 0; it < iterations; ++it) {
    // ping-pong so that the last iteration lands in `out`
    float* dst = ((iterations - 1 - it) & 1) ? ws : out;
    bilateral_fast_iter_kernel<<<tile_grid(N, H, W), kThreads, 0, ST>>>(src, H, W, kernel1d, k, inv, dst);
    HIS_CHECK_LAUNCH();
    src = dst;
  }
  return HIS_OK;
}

int his_post_guided_filter(const float* x, const float* guide, int N, int H, int W, int radius, float eps, float* ws_a, float* ws_b, float* out,
                           void* stream) {
  if (!x || !guide || !ws_a || !ws_b || !out) return his_set_error(HIS_ERR_INVALID_ARG, "guided_filter: null pointer");
  if (radius < 0 || radius > kMaxR) return his_set_error(HIS_ERR_UNSUPPORTED, "guided_filter: radius <= 8");
  CHECK_PLANES("guided_filter");
  guided_ab_kernel<<<tile_grid(N, H, W), kThreads, 0, ST>>>(x, guide, H, W, radius, eps, ws_a, ws_b);
  HIS_CHECK_LAUNCH();
  guided_out_kernel<<<tile_grid(N, H, W), kThreads, 0, ST>>>(ws_a, ws_b, guide, H, W, radius, out);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_post_binary_bilateral_tiled(const float* mask, int N, int H, int W, const float* gauss, int k, int iterations, float threshold, float* out,
                                    void* stream) {
  if (!mask || !gauss || !out) return his_set_error(HIS_ERR_INVALID_ARG, "binary_bilateral_tiled: null pointer");
  if (k < 1 || !(k & 1) || k > 9 || iterations < 1 || iterations * (k / 2) > kMaxR)
    return his_set_error(HIS_ERR_UNSUPPORTED, "binary_bilateral_tiled: odd kernel size <= 9 and iterations*(k/2) <= 8");
  CHECK_PLANES("binary_bilateral_tiled");
  binary_bilateral_smem_kernel<<<tile_grid(N, H, W), kThreads, 0, ST>>>(mask, H, W, gauss, k, iterations, threshold, out);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_post_dilate_logits(const float* logits, int N, int H, int W, int dilation_pixels, float* out, void* stream) {
  if (!logits || !out) return his_set_error(HIS_ERR_INVALID_ARG, "dilate_logits: null pointer");
  if (dilation_pixels < 0 || dilation_pixels > kDilMax) return his_set_error(HIS_ERR_UNSUPPORTED, "dilate_logits: dilation_pixels must be in [0,16]");
  CHECK_PLANES("dilate_logits");
  if (dilation_pixels == 0) {
    if (cudaMemcpyAsync(out, logits, (size_t)N * H * W * 3 * sizeof(float), cudaMemcpyDeviceToDevice, ST) != cudaSuccess)
      return his_set_error(HIS_ERR_LAUNCH, "memcpy failed");
    return HIS_OK;
  }
  launch_dilate<false>(logits, N, H, W, dilation_pixels, 0.0f, out, nullptr, ST);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_post_dilate_instance_mask(const float* logits, int N, int H, int W, int dilation_pixels, float score_threshold, float* out_f32,
                                  unsigned char* out_u8, void* stream) {
  if (!logits || (!out_f32 && !out_u8)) return his_set_error(HIS_ERR_INVALID_ARG, "dilate_instance_mask: null pointer");
  if (dilation_pixels < 0 || dilation_pixels > kDilMax)
    return his_set_error(HIS_ERR_UNSUPPORTED, "dilate_instance_mask: dilation_pixels must be in [0,16]");
  CHECK_PLANES("dilate_instance_mask");
  launch_dilate<true>(logits, N, H, W, dilation_pixels, score_threshold, out_f32, out_u8, ST);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_post_mask_cleanup_fused(const float* mask, int N, int H, int W, float es_threshold, float es_strength, const float* gauss, int k,
                                int iterations, float threshold, float* out, void* stream) {
  if (!mask || !gauss || !out) return his_set_error(HIS_ERR_INVALID_ARG, "mask_cleanup_fused: null pointer");
  if (k < 1 || !(k & 1) || k > 9 || iterations < 1 || 1 + iterations * (k / 2) > kMaxR)
    return his_set_error(HIS_ERR_UNSUPPORTED, "mask_cleanup_fused: odd kernel size <= 9 and 1 + iterations*(k/2) <= 8");
  CHECK_PLANES("mask_cleanup_fused");
  // wide form for the shapes it covers (HIS_POST_WIDE=0 keeps the general form: A/B runs)
  static const bool wide = [] { const char* e = getenv("HIS_POST_WIDE"); return !e || atoi(e) != 0; }();
  if (!wide || !launch_cleanup_wide<float>(mask, N, H, W, es_threshold, es_strength, gauss, k, iterations, threshold, out, ST))
    mask_cleanup_fused_kernel<<<tile_grid(N, H, W), kThreads, 0, ST>>>(mask, H, W, es_threshold, es_strength, gauss, k, iterations, threshold, out);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_post_mask_cleanup_fused_u8(const unsigned char* mask, int N, int H, int W, float es_threshold, float es_strength, const float* gauss,
                                   int k, int iterations, float threshold, unsigned char* out, void* stream) {
  if (!mask || !gauss || !out) return his_set_error(HIS_ERR_INVALID_ARG, "mask_cleanup_fused_u8: null pointer");
  CHECK_PLANES("mask_cleanup_fused_u8");
  if (!launch_cleanup_wide<unsigned char>(mask, N, H, W, es_threshold, es_strength, gauss, k, iterations, threshold, out, ST))
    return his_set_error(HIS_ERR_UNSUPPORTED, "mask_cleanup_fused_u8: kernel size 3 / 5 / 7 and 1 or 2 iterations");
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

}  // extern "C"
