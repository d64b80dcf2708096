// Shared device/host helpers for the sm_100a kernels of the RGB hierarchical
// instance-segmentation path.  Activations live in HBM as NHWC fp16 ("channel-last
// rows"), optionally as a channel slice [c_off, c_off+C) of a wider buffer whose
// per-pixel stride is Cs elements (Cs % 8 == 0 so that TMA strides are 16-byte
// multiples).  Tails with <= 2 channels (logits) are kept in fp32.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define HIS_OK 0
#define HIS_ERR_INVALID_ARG (-1)
#define HIS_ERR_UNSUPPORTED (-2)
#define HIS_ERR_LAUNCH (-3)
#define HIS_ERR_DRIVER (-4)
#define HIS_ERR_NO_DEVICE (-5)

enum HisAct { HIS_ACT_NONE = 0, HIS_ACT_RELU = 1, HIS_ACT_SILU = 2, HIS_ACT_SIGMOID = 3, HIS_ACT_SWISH = 4, HIS_ACT_GELU = 5 };
enum HisResMode { HIS_RES_NONE = 0, HIS_RES_ADD = 1 /* before act */, HIS_RES_MUL = 2 /* after act */ };

#define HIS_CHECK_LAUNCH()                                   \
  do {                                                       \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) return his_set_error(HIS_ERR_LAUNCH, cudaGetErrorString(e__)); \
  } while (0)

int his_set_error(int code, const char* msg);

__device__ __forceinline__ float his_sigmoid(float x) { return 1.0f / (1.0f + __expf(-x)); }

// One MUFU each, no range fix-up code around them (`__expf` / `__fdividef` add an FSETP and two predicated FMULs per value for
// denormal results): 1/(1+2^(nb2*x)) with nb2 = -beta*log2(e).  Flush-to-zero is exact enough here: 2^t < 2^-126 -> s = 1,
// 2^t = inf -> s = 0, and x*s of a finite x stays finite.
__device__ __forceinline__ float his_ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float his_rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float his_sigmoid_fast(float x, float nb2) { return his_rcp_approx(1.0f + his_ex2_approx(nb2 * x)); }
#define HIS_NEG_LOG2E (-1.4426950408889634f)
// packed fp32 FMA of sm_100 (FFMA2): two IEEE fma.rn in one issue slot
__device__ __forceinline__ float2 his_ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}

__device__ __forceinline__ float his_act(float v, int act, float beta) {
  switch (act) {
    case HIS_ACT_RELU: return fmaxf(v, 0.0f);
    case HIS_ACT_SILU: return v * his_sigmoid(v);
    case HIS_ACT_SIGMOID: return his_sigmoid(v);
    case HIS_ACT_SWISH: return v * his_sigmoid(beta * v);
    case HIS_ACT_GELU: return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
    default: return v;
  }
}

static inline int his_div_up(int a, int b) { return (a + b - 1) / b; }

// ---- split-fp16 activations ("strict" precision mode): a value is a pair of fp16 numbers x = hi + lo with hi = fp16(x) and
// lo = fp16(x - hi) (~21 significant bits); the lo plane of a pixel lies `lo` elements after its hi plane inside the same NHWC
// buffer ([hi channels | lo channels], lo = pixel stride / 2).  lo == 0 means plain fp16.  The C entry points take `int split`
// and derive lo = cs / 2 per tensor.
#ifdef __CUDACC__
__device__ __forceinline__ void his_ld8(const __half* p, int lo, float* f) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int e = 0; e < 4; ++e) { const float2 t = __half22float2(h[e]); f[2 * e] = t.x; f[2 * e + 1] = t.y; }
  if (lo) {
    const uint4 w = __ldg(reinterpret_cast<const uint4*>(p + lo));
    const __half2* l = reinterpret_cast<const __half2*>(&w);
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float2 t = __half22float2(l[e]); f[2 * e] += t.x; f[2 * e + 1] += t.y; }
  }
}
__device__ __forceinline__ void his_st8(__half* p, int lo, const float* f) {
  uint4 v;
  __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
  *reinterpret_cast<uint4*>(p) = v;
  if (lo) {
    uint4 w;
    __half2* l = reinterpret_cast<__half2*>(&w);
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float2 t = __half22float2(h[e]); l[e] = __floats2half2_rn(f[2 * e] - t.x, f[2 * e + 1] - t.y); }
    *reinterpret_cast<uint4*>(p + lo) = w;
  }
}
__device__ __forceinline__ float his_ld1(const __half* p, int lo) {
  float v = __half2float(__ldg(p));
  if (lo) v += __half2float(__ldg(p + lo));
  return v;
}
__device__ __forceinline__ void his_st1(__half* p, int lo, float v) {
  const __half h = __float2half_rn(v);
  *p = h;
  if (lo) p[lo] = __float2half_rn(v - __half2float(h));
}
// One-time work per (call site, device): function attributes such as the dynamic shared memory opt-in are per device.
struct PerDeviceOnce {
  bool done[64] = {false};
  bool first() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};

#endif
