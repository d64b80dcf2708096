// Post-processing stencils of the exported pipeline: 3-class argmax ->
// instance mask, MaskDilationModule, BinaryMaskEdgeSmoothing, BinaryMaskBilateralFilter,
// MorphologicalBilateralFilter and NEAREST paste-back.  All HBM-bound: one thread per output
// pixel, rows walked by consecutive lanes (coalesced), neighbourhood re-reads served by L1/L2.
// Float stages keep the reference's operation order (no fast-math) because their final
// thresholds can sit within a few ulp of a tie.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
inline int grid_for(long long work) {
  long long g = (work + kThreads - 1) / kThreads;
  if (g > 148LL * 64) g = 148LL * 64;
  return (int)(g < 1 ? 1 : g);
}
#define GRID_STRIDE(idx, total) \
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < (total); idx += (long long)gridDim.x * blockDim.x)

// hed/export_onnx_advanced.py:360-364: where(argmax(masks,1)==1, 1, 0) (argmax returns the FIRST maximum);
// test_hierarchical_instance_peopleseg_onnx.py:250-262: (argmax(softmax)==1) & (max prob > score_threshold).
__global__ void instance_mask_kernel(const float* __restrict__ logits, int N, long long HW, float thr, float* __restrict__ out_f,
                                     unsigned char* __restrict__ out_u8) {
  const long long total = (long long)N * HW;
  GRID_STRIDE(idx, total) {
    const long long n = idx / HW, p = idx % HW;
    const float l0 = logits[(n * 3) * HW + p], l1 = logits[(n * 3 + 1) * HW + p], l2 = logits[(n * 3 + 2) * HW + p];
    bool on = (l1 > l0) && (l1 >= l2);
    if (on && thr > 0.0f) {
      const float m = l1;
      const float s = expf(l0 - m) + 1.0f + expf(l2 - m);
      on = (1.0f / s) > thr;
    }
    if (out_f) out_f[idx] = on ? 1.0f : 0.0f;
    if (out_u8) out_u8[idx] = on ? 1 : 0;
  }
}

__device__ __forceinline__ float at0(const float* __restrict__ img, int H, int W, int y, int x) {
  return (y >= 0 && y < H && x >= 0 && x < W) ? img[(long long)y * W + x] : 0.0f;
}

// BinaryMaskEdgeSmoothing, hed/edge_smoothing.py:10-90 (per channel == per plane here)
__global__ void edge_smooth_kernel(const float* __restrict__ mask, int N, int H, int W, float thr, float strength, float* __restrict__ out) {
  const long long HW = (long long)H * W, total = (long long)N * HW;
  GRID_STRIDE(idx, total) {
    const long long n = idx / HW, p = idx % HW;
    const int y = (int)(p / W), x = (int)(p % W);
    const float* img = mask + n * HW;
    float v[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) v[t] = at0(img, H, W, y + t / 3 - 1, x + t % 3 - 1);
    float lap = 0.0f, g = 0.0f;
    const float gk[9] = {1.f / 16, 2.f / 16, 1.f / 16, 2.f / 16, 4.f / 16, 2.f / 16, 1.f / 16, 2.f / 16, 1.f / 16};
#pragma unroll
    for (int t = 0; t < 9; ++t) { lap = fmaf(v[t], t == 4 ? 8.0f : -1.0f, lap); g = fmaf(v[t], gk[t], g); }
    const float e = fabsf(lap) * strength;
    const float w = 1.0f / (1.0f + expf(-e));
    const float sm = __fadd_rn(__fmul_rn(v[4], __fsub_rn(1.0f, w)), __fmul_rn(g, w));
    out[idx] = sm > thr ? 1.0f : 0.0f;
  }
}

// one iteration of BinaryMaskBilateralFilter (hed/bilateral_filter.py:299-406); `first` clamps the input to [0,1]
__global__ void binary_bilateral_iter_kernel(const float* __restrict__ in, int N, int H, int W, const float* __restrict__ gk, int k, int clamp_in,
                                             int last, float thr, float* __restrict__ out) {
  const long long HW = (long long)H * W, total = (long long)N * HW;
  const int r = k / 2;
  GRID_STRIDE(idx, total) {
    const long long n = idx / HW, p = idx % HW;
    const int y = (int)(p / W), x = (int)(p % W);
    const float* img = in + n * HW;
    float f = 0.0f, f2 = 0.0f;
    for (int ky = 0; ky < k; ++ky) {
      const int yy = y + ky - r;
      if (yy < 0 || yy >= H) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int xx = x + kx - r;
        if (xx < 0 || xx >= W) continue;
        float v = img[(long long)yy * W + xx];
        if (clamp_in) v = fminf(fmaxf(v, 0.0f), 1.0f);
        const float wgt = __ldg(gk + ky * k + kx);
        f = fmaf(v, wgt, f);
        f2 = fmaf(__fmul_rn(v, v), wgt, f2);
      }
    }
    float c = img[p];
    if (clamp_in) c = fminf(fmaxf(c, 0.0f), 1.0f);
    const float var = fmaxf(__fsub_rn(f2, __fmul_rn(f, f)), 0.0f);
    const float ew = expf(__fmul_rn(-var, 10.0f));
    const float m = __fadd_rn(__fmul_rn(ew, f), __fmul_rn(__fsub_rn(1.0f, ew), c));
    out[idx] = last ? (m > thr ? 1.0f : 0.0f) : m;
  }
}

// window max / min with the window clipped at the border (max_pool2d pads with -inf; erosion = -maxpool(-x))
__global__ void minmax_kernel(const float* __restrict__ in, int N, int H, int W, int k, int is_max, int clamp_in, float* __restrict__ out) {
  const long long HW = (long long)H * W, total = (long long)N * HW;
  const int r = k / 2;
  GRID_STRIDE(idx, total) {
    const long long n = idx / HW, p = idx % HW;
    const int y = (int)(p / W), x = (int)(p % W);
    const float* img = in + n * HW;
    float m = is_max ? -INFINITY : INFINITY;
    for (int dy = -r; dy <= r; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= H) continue;
      for (int dx = -r; dx <= r; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= W) continue;
        float v = img[(long long)yy * W + xx];
        if (clamp_in) v = fminf(fmaxf(v, 0.0f), 1.0f);
        m = is_max ? fmaxf(m, v) : fminf(m, v);
      }
    }
    out[idx] = m;
  }
}

__global__ void conv_zero_pad_kernel(const float* __restrict__ in, int N, int H, int W, const float* __restrict__ kern, int k, float* __restrict__ out) {
  const long long HW = (long long)H * W, total = (long long)N * HW;
  const int r = k / 2;
  GRID_STRIDE(idx, total) {
    const long long n = idx / HW, p = idx % HW;
    const int y = (int)(p / W), x = (int)(p % W);
    const float* img = in + n * HW;
    float f = 0.0f;
    for (int ky = 0; ky < k; ++ky) {
      const int yy = y + ky - r;
      if (yy < 0 || yy >= H) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int xx = x + kx - r;
        if (xx < 0 || xx >= W) continue;
        f = fmaf(img[(long long)yy * W + xx], __ldg(kern + ky * k + kx), f);
      }
    }
    out[idx] = f;
  }
}

__global__ void threshold_kernel(const float* __restrict__ in, long long total, float thr, float* __restrict__ out) {
  GRID_STRIDE(idx, total) out[idx] = in[idx] > thr ? 1.0f : 0.0f;
}

// Paste-back (test_hierarchical_instance_peopleseg_onnx.py:144-161,264-278,369-374): box = int(x1*W) ... (fp32 product,
// truncation), cv2.resize(..., INTER_NEAREST): src = min(floor(dst * (1/(dst_size/src_size))), src_size-1) in double,
// full[y1:y2, x1:x2] = mask; later instances win.  canvas[b,y,x] = 1 + index of the last ROI whose pasted mask is 1.
// One CTA = one ROI x one slice of its rows: the source column of every destination column is tabulated once per CTA (the
// double-precision floor product of cv2's NEAREST rule), a warp then walks one destination row at a time (source row computed
// once per row, lanes over columns: coalesced RED.MAX, byte gather along one mask row) -- no 64-bit division per pixel.
constexpr int kPasteTab = 4096;
__global__ void __launch_bounds__(kThreads) paste_kernel(const unsigned char* __restrict__ masks, int N, int mh, int mw, const float* __restrict__ rois,
                                                         int* __restrict__ canvas, int B, int H, int W) {
  __shared__ unsigned short sxt[kPasteTab];
  const int roi = blockIdx.x;          // ROIs on gridDim.x (2^31-1 blocks): one call handles any ROI count
  const float* r = rois + 5 * roi;
  const int b = (int)r[0];
  if (b < 0 || b >= B) return;
  const int x1 = (int)__fmul_rn(r[1], (float)W), y1 = (int)__fmul_rn(r[2], (float)H);
  const int x2 = (int)__fmul_rn(r[3], (float)W), y2 = (int)__fmul_rn(r[4], (float)H);
  const int bw = x2 - x1, bh = y2 - y1;
  if (bw <= 0 || bh <= 0) return;
  const double sx = 1.0 / ((double)bw / (double)mw), sy = 1.0 / ((double)bh / (double)mh);
  const bool tab = bw <= kPasteTab && mw <= 65536;
  if (tab) {
    for (int dx = threadIdx.x; dx < bw; dx += kThreads) sxt[dx] = (unsigned short)min((int)floor((double)dx * sx), mw - 1);
    __syncthreads();
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = kThreads / 32;
  // destination rows / columns clipped to the canvas once
  const int dy_lo = max(0, -y1), dy_hi = min(bh, H - y1), dx_lo = max(0, -x1), dx_hi = min(bw, W - x1);
  for (int dy = dy_lo + blockIdx.y * wpb + warp; dy < dy_hi; dy += gridDim.y * wpb) {
    const int syi = min((int)floor((double)dy * sy), mh - 1);
    const unsigned char* mrow = masks + ((long long)roi * mh + syi) * mw;
    int* crow = canvas + ((long long)b * H + (y1 + dy)) * W + x1;
    if (tab) {                                              // two loops: a select would keep the double-precision product in the hot one
      for (int dx = dx_lo + lane; dx < dx_hi; dx += 32)
        if (mrow[sxt[dx]]) atomicMax(crow + dx, roi + 1);
    } else {
      for (int dx = dx_lo + lane; dx < dx_hi; dx += 32)
        if (mrow[min((int)floor((double)dx * sx), mw - 1)]) atomicMax(crow + dx, roi + 1);
    }
  }
}

// Input path (test_hierarchical_instance_peopleseg_onnx.py:170-196, after the file decode): BGR->RGB, cv2.resize(uint8,
// INTER_LINEAR), /255, HWC -> CHW, fused.  OpenCV's 8-bit bilinear arithmetic is kept bit for bit: 11-bit fixed-point weights
// (tables built by the caller exactly like cv2 builds them), horizontal pass in int32, vertical pass
// (((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2.  One thread = one output pixel, three channels.
__global__ void preprocess_u8_kernel(const unsigned char* __restrict__ src, int N, int Hs, int Ws, int Hd, int Wd, const int* __restrict__ xtab,
                                     const int* __restrict__ ytab, int swap_rb, float* __restrict__ out) {
  const long long total = (long long)N * Hd * Wd;
  GRID_STRIDE(idx, total) {
    const int x = (int)(idx % Wd), y = (int)((idx / Wd) % Hd);
    const long long n = idx / ((long long)Wd * Hd);
    const int x0 = xtab[x], x1 = xtab[Wd + x], ax0 = xtab[2 * Wd + x], ax1 = xtab[3 * Wd + x];
    const int y0 = ytab[y], y1 = ytab[Hd + y], ay0 = ytab[2 * Hd + y], ay1 = ytab[3 * Hd + y];
    const unsigned char* img = src + n * (long long)Hs * Ws * 3;
    const unsigned char* r0 = img + (long long)y0 * Ws * 3;
    const unsigned char* r1 = img + (long long)y1 * Ws * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int sc = swap_rb ? 2 - c : c;
      const int h0 = (int)r0[x0 * 3 + sc] * ax0 + (int)r0[x1 * 3 + sc] * ax1;
      const int h1 = (int)r1[x0 * 3 + sc] * ax0 + (int)r1[x1 * 3 + sc] * ax1;
      const int v = (((ay0 * (h0 >> 4)) >> 16) + ((ay1 * (h1 >> 4)) >> 16) + 2) >> 2;
      const int u = min(max(v, 0), 255);
      out[((n * 3 + c) * Hd + y) * (long long)Wd + x] = __fdiv_rn((float)u, 255.0f);
    }
  }
}

// evaluate_model's metric core (hed/train_utils.py:262-292): per-ROI 3x3 confusion counts of (ground-truth class, argmax
// class).  The reference moves predictions to the CPU and runs Python double loops per sample; everything it reports --
// the three confusion matrices, per-class IoUs, detection rates -- is a function of these nine integers per ROI.
// One CTA per (ROI, slice): warp-shuffle + shared-memory integer reduction, then one atomicAdd per counter (integers:
// order independent, deterministic).  gt: class labels 0..2 as uint8 or int64.
template <typename T>
__global__ void eval_confusion_kernel(const float* __restrict__ logits, const T* __restrict__ gt, long long HW, int* __restrict__ counts) {
  __shared__ int s_cnt[9];
  const int n = blockIdx.x;            // ROIs on gridDim.x
  if (threadIdx.x < 9) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  int c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  const float* l = logits + (long long)n * 3 * HW;
  for (long long p = blockIdx.y * (long long)blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.y * blockDim.x) {
    const float l0 = l[p], l1 = l[HW + p], l2 = l[2 * HW + p];
    const int pred = (l1 > l0) ? ((l2 > l1) ? 2 : 1) : ((l2 > l0) ? 2 : 0);      // torch.argmax: first maximum
    const int t = (int)gt[(long long)n * HW + p];
    if (t >= 0 && t < 3) {
#pragma unroll
      for (int k = 0; k < 9; ++k) c[k] += (k == t * 3 + pred) ? 1 : 0;
    }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    int v = c[k];
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_cnt[k], v);
  }
  __syncthreads();
  if (threadIdx.x < 9 && s_cnt[threadIdx.x]) atomicAdd(counts + n * 9 + threadIdx.x, s_cnt[threadIdx.x]);
}

}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" {

int his_post_instance_mask(const float* logits, int N, int H, int W, float score_threshold, float* out_f32, unsigned char* out_u8, void* stream) {
  if (!logits || (!out_f32 && !out_u8)) return his_set_error(HIS_ERR_INVALID_ARG, "instance_mask: null pointer");
  const long long total = (long long)N * H * W;
  if (total == 0) return HIS_OK;
  instance_mask_kernel<<<grid_for(total), kThreads, 0, ST>>>(logits, N, (long long)H * W, score_threshold, out_f32, out_u8);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_post_edge_smooth(const float* mask, int N, int H, int W, float threshold, float blur_strength, float* out, void* stream) {
  if (!mask || !out) return his_set_error(HIS_ERR_INVALID_ARG, "edge_smooth: null pointer");
  const long long total = (long long)N * H * W;
  if (total == 0) return HIS_OK;
  edge_smooth_kernel<<<grid_for(total), kThreads, 0, ST>>>(mask, N, H, W, threshold, blur_strength, out);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_post_binary_bilateral(const float* mask, int N, int H, int W, const float* gauss, int k, int iterations, float threshold, float* ws0,
                              float* ws1, float* out, void* stream) {
  if (!mask || !gauss || !out || (iterations > 1 && (!ws0 || !ws1))) return his_set_error(HIS_ERR_INVALID_ARG, "binary_bilateral: null pointer");
  if (k < 1 || !(k & 1) || iterations < 1) return his_set_error(HIS_ERR_INVALID_ARG, "binary_bilateral: kernel size must be odd, iterations >= 1");
  const long long total = (long long)N * H * W;
  if (total == 0) return HIS_OK;
  const float* src = mask;
  for (int it = 0; it < iterations; ++it) {
    const int last = it == iterations - 1;
    float* dst = last ? out : (it & 1 ? ws1 : ws0);
    binary_bilateral_iter_kernel<<<grid_for(total), kThreads, 0, ST>>>(src, N, H, W, gauss, k, it == 0, last, threshold, dst);
    HIS_CHECK_LAUNCH();
    src = dst;
  }
  return HIS_OK;
}

int his_post_morph_bilateral(const float* mask, int N, int H, int W, const float* kernel2d, int k, int morph, float* ws0, float* ws1, float* out,
                             void* stream) {
  if (!mask || !kernel2d || !ws0 || !ws1 || !out) return his_set_error(HIS_ERR_INVALID_ARG, "morph_bilateral: null pointer");
  const long long total = (long long)N * H * W;
  if (total == 0) return HIS_OK;
  const int g = grid_for(total);
  minmax_kernel<<<g, kThreads, 0, ST>>>(mask, N, H, W, morph, 0, 1, ws0);        // open: erode(clamp(x))
  minmax_kernel<<<g, kThreads, 0, ST>>>(ws0, N, H, W, morph, 1, 0, ws1);         //       dilate
  conv_zero_pad_kernel<<<g, kThreads, 0, ST>>>(ws1, N, H, W, kernel2d, k, ws0);  // gaussian, zero padded
  minmax_kernel<<<g, kThreads, 0, ST>>>(ws0, N, H, W, morph, 1, 0, ws1);         // close: dilate
  minmax_kernel<<<g, kThreads, 0, ST>>>(ws1, N, H, W, morph, 0, 0, ws0);         //        erode
  threshold_kernel<<<g, kThreads, 0, ST>>>(ws0, total, 0.5f, out);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_preprocess_u8(const unsigned char* src, int N, int Hs, int Ws, int Hd, int Wd, const int* xtab, const int* ytab, int swap_rb, float* out,
                      void* stream) {
  if (!src || !xtab || !ytab || !out) return his_set_error(HIS_ERR_INVALID_ARG, "preprocess_u8: null pointer");
  const long long total = (long long)N * Hd * Wd;
  if (total == 0) return HIS_OK;
  if (Hs <= 0 || Ws <= 0) return his_set_error(HIS_ERR_INVALID_ARG, "preprocess_u8: empty source image");
  preprocess_u8_kernel<<<grid_for(total), kThreads, 0, ST>>>(src, N, Hs, Ws, Hd, Wd, xtab, ytab, swap_rb, out);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_eval_confusion(const float* logits, const void* gt, int gt_is_int64, int N, int H, int W, int* counts, void* stream) {
  if (!logits || !gt || !counts) return his_set_error(HIS_ERR_INVALID_ARG, "eval_confusion: null pointer");
  if (N == 0) return HIS_OK;
  if (cudaMemsetAsync(counts, 0, (size_t)N * 9 * sizeof(int), ST) != cudaSuccess) return his_set_error(HIS_ERR_LAUNCH, "memset failed");
  const long long HW = (long long)H * W;
  int gx = (int)((HW + kThreads * 8 - 1) / (kThreads * 8));
  if (gx < 1) gx = 1;
  if (gx > 65535) gx = 65535;
  dim3 grid(N, gx);
  if (gt_is_int64) eval_confusion_kernel<long long><<<grid, kThreads, 0, ST>>>(logits, (const long long*)gt, HW, counts);
  else eval_confusion_kernel<unsigned char><<<grid, kThreads, 0, ST>>>(logits, (const unsigned char*)gt, HW, counts);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_post_paste(const unsigned char* masks, int N, int mh, int mw, const float* rois, int* canvas, int B, int H, int W, void* stream) {
  if (!masks || !rois || !canvas) return his_set_error(HIS_ERR_INVALID_ARG, "paste: null pointer");
  if (N == 0) return HIS_OK;
  dim3 grid(N, 8);
  paste_kernel<<<grid, kThreads, 0, ST>>>(masks, N, mh, mw, rois, canvas, B, H, W);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

}  // extern "C"
