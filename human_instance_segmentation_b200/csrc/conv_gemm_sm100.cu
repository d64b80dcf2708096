// Implicit-GEMM convolution on the Blackwell tensor cores (tcgen05 + TMEM), fed by TMA.
//
// Replaces the dense conv2d / conv_transpose2d(k2,s2) ATen calls of the reference head
// (hed/advanced/hierarchical_segmentation_rgb.py:657-673, ..._refinement.py:37-39,479-523,
// ..._unet.py:44-47,313-372) and of the EfficientNet-UNet 1x1 / decoder 3x3 convs.
//
// GEMM view: D[M=128 pixels, N=Cout tile] += A[M, K] * B[N, K]^T, K = taps * Cin.
//  * A is never materialised: the M tile is a (bh x bw) rectangle of output pixels of one image
//    (bh*bw == 128) and, for filter tap (dy,dx), the operand tile is the SAME rectangle of the
//    NHWC fp16 input shifted by (dy-1,dx-1).  One 4-D TMA box load {64ch, bw, bh, 1} per
//    (tap, 64-channel block) lands it in shared memory as 128 rows x 128 B in the canonical
//    K-major SWIZZLE_128B UMMA layout; out-of-image coordinates are zero-filled by the TMA
//    unit, which *is* the conv zero padding (and the channel tail padding).
//  * B (weights) is pre-packed fp16 [group][tap][Cout_slab][Cin_pad], K-major, 2-D TMA tiles.
//  * D accumulates in TMEM (fp32), double buffered (2 x block_n columns) so that the epilogue
//    of tile i overlaps the MMAs of tile i+1.
//  * Epilogue: tcgen05.ld -> y = act(acc + shift[c] (+res)) (*res) -> fp16 -> swizzled smem staging ->
//    TMA store (clips partial tiles / channel tails), 32 channels at a time.  The BatchNorm scale is
//    folded into the packed weights by the caller; the residual / multiplicand tile arrives through
//    TMA into the same staging buffer.  Two epilogue warpgroups alternate tiles (one per TMEM
//    accumulator), so two tiles drain concurrently while the MMAs of later tiles run.
//  * conv_transpose k2s2 = 4 independent 1x1 GEMMs ("groups"), each scattered through its own
//    strided output tensor map (pixel (2y+dy, 2x+dx)).
//
// Warp roles (384 threads, 1 CTA/SM, persistent over work items):
//   warp 0: TMA producer   warp 1: MMA issuer   warp 2: TMEM allocator   warps 4-7 / 8-11: epilogue groups 0 / 1.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>

#include "common.cuh"

namespace {

constexpr int kBlockM = 128;
constexpr int kChunkC = 32;                            // output channels per epilogue chunk
constexpr int kStagingBytes = kBlockM * kChunkC * 2;   // 8 KB : 128 rows x 32 channels fp16 (SWIZZLE_64B rows)
constexpr int kNumStaging = 4;                         // two per epilogue group
constexpr int kShiftBytes = 256 * 4;                   // per-channel shift of the (single) N tile
constexpr int kThreadsGemm = 384;
constexpr int kTmemCols = 512;
constexpr int kMaxStages = 48;
constexpr int kSmemBudget = 227 * 1024;
constexpr int kBarrierBytes = 1024;

// K-block width BK (fp16 elements) selects the shared-memory swizzle: one row of the operand tile is BK*2 bytes.
template <int BK> struct KCfg {
  static constexpr int kABytes = kBlockM * BK * 2;
  static constexpr int kBBytesMax = 256 * BK * 2;
  // the ring gets whatever the 227 KB leave after the staging buffers, barriers and the 1 KB alignment slack; its depth is a
  // run-time parameter (stage = A + the layer's actual B tile), so narrow layers keep many more loads in flight
  static constexpr int kRingBytes = kSmemBudget - kNumStaging * kStagingBytes - kShiftBytes - 1024 - kBarrierBytes;
  static constexpr int kSmemBytes = kSmemBudget;
  static constexpr uint32_t kSbo = 8 * BK * 2;                 // bytes between 8-row groups
  static constexpr uint64_t kLayout = BK == 64 ? 2 : BK == 32 ? 4 : 6;   // SWIZZLE_128B / 64B / 32B
};

enum { ACTC_CLAMP = 0, ACTC_SIGMOID = 1, ACTC_GELU = 2 };

struct ConvGemmParams {
  int n_img, H, W;
  int bh, bw, tiles_x, tiles_y;
  int ksize;            // 1 or 3
  int kblocks_per_tap;  // ceil(Cin / 64)
  int n_tiles, block_n; // N tiling of one (group, tap) slab
  int groups;           // 1, or 4 for conv-transpose k2s2
  int cout_slab;        // n_tiles * block_n  (rows of one slab in B, scale/shift length per group)
  int num_work;
  int b_img_rows;       // rows to skip in B per image (0: shared weights; >0: per-image weights, e.g. SE gate folded in)
  int stages, stage_bytes;   // smem ring: stage = A tile (128 x BK) + B tile (block_n x BK), rounded up to 1 KB
  // activation, compile-time class + runtime parameters:
  //   CLAMP:   y = max(y, act_lo)                  (none: -inf, relu: 0)
  //   SIGMOID: s = 1/(1+exp(-act_beta*y)); y = act_mul_x ? y*s : s   (sigmoid / silu / swish(beta))
  //   GELU:    exact erf form
  float act_lo, act_beta;
  int act_mul_x;
  const float* shift;   // [cout_slab]  (conv bias + folded-BatchNorm shift; the BN scale lives in the weights)
  // fused 1x1 tail to <= 2 channels (TAIL kernels): tail[o] = sum_c y[c]*tail_w[o][c] + tail_b[o], NCHW fp32 out
  const float* tail_w;  // [tail_c][cout_slab]
  float* tail_out;      // [n_img][tail_c][H][W]
  float tail_b0, tail_b1;
  int tail_c, tail_sigmoid, store_main;
  // fp32 NCHW copy of the layer output (EPI_AUX kernels): aux[n][c][y][x] = y, or the gate itself (before the product) for RES_MUL
  float* aux_out;
  int cout;             // true output channel count
};

// ------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// one lane of the (converged) warp; keeps the surrounding code warp-uniform so that the uniform-datapath instructions
// (UTMALDG / UTCHMMA / UTCBAR) are issued directly instead of through a per-lane election loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}

// K-major swizzled shared-memory matrix descriptor (sm_100 format), one swizzle atom along K (row = BK*2 bytes):
//   [0,14) start>>4 | [16,30) LBO>>4 (unused) | [32,46) SBO>>4 (8 rows) | [46,48) version=1 | [61,64) swizzle mode
template <int BK>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(KCfg<BK>::kSbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= KCfg<BK>::kLayout << 61;
  return d;
}

// kind::f16 instruction descriptor: D=f32, A=B=f16, both K-major, M=128, N=n.
__device__ __forceinline__ uint32_t make_idesc_f16(int n) {
  uint32_t d = 0;
  d |= 1u << 4;                       // c_format = F32
  d |= 0u << 7;                       // a_format = F16
  d |= 0u << 10;                      // b_format = F16
  d |= (uint32_t)(n >> 3) << 17;      // n_dim
  d |= (uint32_t)(kBlockM >> 4) << 24;  // m_dim
  return d;
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int ACTC>
__device__ __forceinline__ float epi_act(float y, const ConvGemmParams& p) {
  if (ACTC == ACTC_CLAMP) return fmaxf(y, p.act_lo);
  if (ACTC == ACTC_SIGMOID) {
    const float s = __fdividef(1.0f, 1.0f + __expf(-p.act_beta * y));
    return p.act_mul_x ? y * s : s;
  }
  return 0.5f * y * (1.0f + erff(y * 0.70710678118654752f));
}

struct WorkItem { int img, y0, x0, n_tile, group; };

__device__ __forceinline__ WorkItem decode_work(const ConvGemmParams& p, int w) {
  WorkItem it;
  it.n_tile = w % p.n_tiles; w /= p.n_tiles;
  it.group = w % p.groups;   w /= p.groups;
  int tx = w % p.tiles_x;    w /= p.tiles_x;
  int ty = w % p.tiles_y;    w /= p.tiles_y;
  it.img = w; it.y0 = ty * p.bh; it.x0 = tx * p.bw;
  return it;
}

// ------------------------------------------------------------------------------------ kernel
enum { EPI_PLAIN = 0, EPI_TAIL = 1, EPI_AUX = 2 };

template <int BK, int ACTC, int RES, int EPI>
__global__ void __launch_bounds__(kThreadsGemm, 1)
conv_gemm_sm100_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1,
                       const __grid_constant__ CUtensorMap tmO2, const __grid_constant__ CUtensorMap tmO3,
                       const __grid_constant__ CUtensorMap tmR, const ConvGemmParams p) {
  using Cfg = KCfg<BK>;
  const int kStages = p.stages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_base = smem_base;
  const uint32_t staging_base = smem_base + kStages * p.stage_bytes;
  const uint32_t shift_base = staging_base + kNumStaging * kStagingBytes;
  const uint32_t bar_base = shift_base + kShiftBytes;
  // barrier slots (8 B each): full[kStages] empty[kStages] tmem_full[2] tmem_empty[2] res_full[4]; then the TMEM pointer
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  auto res_bar = [&](int b) { return bar_base + 8u * (2 * kStages + 4 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 8);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));   // generic pointer to the aligned base

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int taps = p.ksize * p.ksize;
  const int kiters = taps * p.kblocks_per_tap;
  const uint32_t b_bytes = (uint32_t)p.block_n * BK * 2;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB); prefetch_tmap(&tmO0);
    if (RES) prefetch_tmap(&tmR);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
    for (int a = 0; a < 4; ++a) mbar_init(res_bar(a), 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // single N tile: the per-channel shift stays in shared memory for the whole kernel
  float* s_shift = reinterpret_cast<float*>(smem_gen + (shift_base - smem_base));
  if (p.n_tiles == 1)
    for (int i = threadIdx.x; i < p.block_n; i += kThreadsGemm) s_shift[i] = __ldg(p.shift + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  if (warp == 0) {
    // ================================ TMA producer (whole warp converged, one elected lane issues) ================================
    int stage = 0; uint32_t phase = 0;
    const int pad = p.ksize >> 1;
    for (int w = blockIdx.x; w < p.num_work; w += gridDim.x) {
      const WorkItem it = decode_work(p, w);
      int brow = it.group * taps * p.cout_slab + it.n_tile * p.block_n + it.img * p.b_img_rows;
      int dy = -pad, dx = -pad, cb = 0;
      for (int k = 0; k < kiters; ++k) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        if (elect_one()) {
          const uint32_t sa = stage_base + stage * p.stage_bytes, sb = sa + Cfg::kABytes;
          mbar_expect_tx(full_bar(stage), Cfg::kABytes + b_bytes);
          tma_load_4d(sa, &tmA, full_bar(stage), cb * BK, it.x0 + dx, it.y0 + dy, it.img);
          tma_load_2d(sb, &tmB, full_bar(stage), cb * BK, brow);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
        if (++cb == p.kblocks_per_tap) { cb = 0; brow += p.cout_slab; if (++dx > pad) { dx = -pad; ++dy; } }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (whole warp converged, one elected lane issues) ================================
    const uint32_t idesc = make_idesc_f16(p.block_n);
    int stage = 0; uint32_t phase = 0; int iter = 0;
    for (int w = blockIdx.x; w < p.num_work; w += gridDim.x, ++iter) {
      const int acc = iter & 1; const uint32_t acc_phase = (iter >> 1) & 1;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
      for (int k = 0; k < kiters; ++k) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = stage_base + stage * p.stage_bytes, sb = sa + Cfg::kABytes;
          const uint64_t adesc = make_kmajor_desc<BK>(sa), bdesc = make_kmajor_desc<BK>(sb);
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk)
            umma_f16(d_tmem, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (k | kk) ? 1u : 0u);
          umma_commit(empty_bar(stage));      // frees the smem slot when these MMAs retire
          if (k == kiters - 1) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue (two groups of 4 warps, group g drains accumulator g) ================================
    const int g = (warp - 4) >> 2;
    const int te = (threadIdx.x - 128) & 127;   // 0..127 == output row of the tile == TMEM lane
    const int q = warp & 3;                     // TMEM lane quarter this warp may touch
    const bool issuer_warp = q == 0;            // first warp of the group issues the group's TMA traffic
    const int nchunks = (p.block_n + kChunkC - 1) / kChunkC;
    const uint32_t stg0 = staging_base + g * 2 * kStagingBytes;
    const uint32_t acc_col = (uint32_t)(g * 256);
    uint32_t cc = 0;
    int iter_g = 0;
    for (int w = blockIdx.x + g * gridDim.x; w < p.num_work; w += 2 * gridDim.x, ++iter_g) {
      const WorkItem it = decode_work(p, w);
      const uint32_t acc_phase = iter_g & 1;
      const CUtensorMap* tmO = it.group == 0 ? &tmO0 : it.group == 1 ? &tmO1 : it.group == 2 ? &tmO2 : &tmO3;
      const int chbase = it.n_tile * p.block_n;
      const float* shp = p.n_tiles == 1 ? s_shift : p.shift + chbase;
      if (RES && issuer_warp && elect_one()) {  // prefetch the first two residual chunks while the MMAs run
        tma_wait_read<0>();
        for (int j = 0; j < 2 && j < nchunks; ++j) {
          const int b = (cc + j) & 1;
          mbar_expect_tx(res_bar(2 * g + b), kStagingBytes);
          tma_load_4d(stg0 + b * kStagingBytes, &tmR, res_bar(2 * g + b), chbase + j * kChunkC, it.x0, it.y0, it.img);
        }
      }
      mbar_wait(tfull_bar(g), acc_phase);
      tc_fence_after();
      float tacc0 = 0.0f, tacc1 = 0.0f;
      constexpr bool TAIL = EPI == EPI_TAIL;
      const bool store_main = !TAIL || p.store_main;
      const int py = it.y0 + te / p.bw, px = it.x0 + te % p.bw;
      float* aux_px = nullptr;
      if (EPI == EPI_AUX && py < p.H && px < p.W) aux_px = p.aux_out + ((long long)it.img * p.cout * p.H + py) * p.W + px;
      for (int j = 0; j < nchunks; ++j, ++cc) {
        const int b = cc & 1;
        const uint32_t stg = stg0 + b * kStagingBytes;
        const int cl0 = j * kChunkC;            // first channel of this chunk within the N tile
        const int ch0 = chbase + cl0;           // ... and within the layer
        const int ncol = min(kChunkC, p.block_n - cl0);   // valid accumulator columns in this chunk (16 or 32)
        if ((!RES || j >= 2) && issuer_warp && elect_one()) {
          tma_wait_read<1>();                    // the store that last read staging[b] has drained
          if (RES) {
            mbar_expect_tx(res_bar(2 * g + b), kStagingBytes);
            tma_load_4d(stg, &tmR, res_bar(2 * g + b), ch0, it.x0, it.y0, it.img);
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
        uint32_t v[kChunkC];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc_col + (uint32_t)cl0;
        if (ncol > 16) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
        tmem_ld_wait();
        if (j == nchunks - 1) {                  // accumulator fully read -> hand TMEM back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(g));
        }
        if (RES) mbar_wait(res_bar(2 * g + b), (cc >> 1) & 1);
        uint8_t* row_ptr = smem_gen + (stg - smem_base) + te * (kChunkC * 2);
#pragma unroll
        for (int i = 0; i < kChunkC / 8; ++i) {
          if (i * 8 < ncol) {
            const int cl = cl0 + i * 8, c = ch0 + i * 8;
            uint4* cell = reinterpret_cast<uint4*>(row_ptr + ((i ^ ((te >> 1) & 3)) << 4));   // SWIZZLE_64B
            float r[8];
            if (RES) {
              const uint4 rv = *cell;
              const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
              for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(rh[e]); r[2 * e] = f.x; r[2 * e + 1] = f.y; }
            }
            const float4 t0 = *reinterpret_cast<const float4*>(shp + cl), t1 = *reinterpret_cast<const float4*>(shp + cl + 4);
            const float sh[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
            float y[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float t = __uint_as_float(v[i * 8 + e]) + sh[e];
              if (RES == HIS_RES_ADD) t += r[e];
              t = epi_act<ACTC>(t, p);
              if (EPI == EPI_AUX) { if (aux_px && c + e < p.cout) aux_px[(long long)(c + e) * p.H * p.W] = t; }
              if (RES == HIS_RES_MUL) t *= r[e];
              y[e] = t;
            }
            if (TAIL) {
              const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.tail_w + c)), w1 = __ldg(reinterpret_cast<const float4*>(p.tail_w + c + 4));
              tacc0 += y[0] * w0.x + y[1] * w0.y + y[2] * w0.z + y[3] * w0.w + y[4] * w1.x + y[5] * w1.y + y[6] * w1.z + y[7] * w1.w;
              if (p.tail_c > 1) {
                const float4 u0 = __ldg(reinterpret_cast<const float4*>(p.tail_w + p.cout_slab + c));
                const float4 u1 = __ldg(reinterpret_cast<const float4*>(p.tail_w + p.cout_slab + c + 4));
                tacc1 += y[0] * u0.x + y[1] * u0.y + y[2] * u0.z + y[3] * u0.w + y[4] * u1.x + y[5] * u1.y + y[6] * u1.z + y[7] * u1.w;
              }
            }
            if (store_main) {
              __half2 o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) o[e] = __floats2half2_rn(y[2 * e], y[2 * e + 1]);
              *cell = *reinterpret_cast<uint4*>(o);
            }
          }
        }
        if (store_main) {
          fence_proxy_async();
          asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
          if (issuer_warp && elect_one()) {
            tma_store_4d(tmO, stg, ch0, it.x0, it.y0, it.img);
            tma_commit();
          }
        }
      }
      if (TAIL) {
        if (py < p.H && px < p.W) {
          float o0 = tacc0 + p.tail_b0, o1 = tacc1 + p.tail_b1;
          if (p.tail_sigmoid) { o0 = 1.0f / (1.0f + __expf(-o0)); o1 = 1.0f / (1.0f + __expf(-o1)); }
          float* dst = p.tail_out + ((long long)it.img * p.tail_c * p.H + py) * p.W + px;
          dst[0] = o0;
          if (p.tail_c > 1) dst[(long long)p.H * p.W] = o1;
        }
      }
    }
    if (issuer_warp && elect_one()) tma_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

typedef void (*ConvGemmKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                               const CUtensorMap, const CUtensorMap, const ConvGemmParams);

template <int BK, int ACTC>
ConvGemmKernel pick_res(int res) {
  return res == 0 ? conv_gemm_sm100_kernel<BK, ACTC, 0, EPI_PLAIN> : res == 1 ? conv_gemm_sm100_kernel<BK, ACTC, 1, EPI_PLAIN>
                                                                            : conv_gemm_sm100_kernel<BK, ACTC, 2, EPI_PLAIN>;
}
template <int BK>
ConvGemmKernel pick_act(int actc, int res) {
  return actc == 0 ? pick_res<BK, 0>(res) : actc == 1 ? pick_res<BK, 1>(res) : pick_res<BK, 2>(res);
}
ConvGemmKernel pick_kernel(int bk, int actc, int res) {
  return bk == 64 ? pick_act<64>(actc, res) : bk == 32 ? pick_act<32>(actc, res) : pick_act<16>(actc, res);
}
// fused-tail variants exist for the shapes that need them: clamp activations (none/relu), no residual or residual-add
ConvGemmKernel pick_tail_kernel(int bk, int res) {
  if (bk == 64) return res == 0 ? conv_gemm_sm100_kernel<64, ACTC_CLAMP, 0, EPI_TAIL> : conv_gemm_sm100_kernel<64, ACTC_CLAMP, 1, EPI_TAIL>;
  if (bk == 32) return res == 0 ? conv_gemm_sm100_kernel<32, ACTC_CLAMP, 0, EPI_TAIL> : conv_gemm_sm100_kernel<32, ACTC_CLAMP, 1, EPI_TAIL>;
  return res == 0 ? conv_gemm_sm100_kernel<16, ACTC_CLAMP, 0, EPI_TAIL> : conv_gemm_sm100_kernel<16, ACTC_CLAMP, 1, EPI_TAIL>;
}
// fp32 NCHW export variants: BK = 64 layers only (the 256-channel trunk / gate layers whose outputs the reference returns)
template <int ACTC>
ConvGemmKernel pick_aux_res(int res) {
  return res == 0 ? conv_gemm_sm100_kernel<64, ACTC, 0, EPI_AUX> : res == 1 ? conv_gemm_sm100_kernel<64, ACTC, 1, EPI_AUX>
                                                                           : conv_gemm_sm100_kernel<64, ACTC, 2, EPI_AUX>;
}
ConvGemmKernel pick_aux_kernel(int actc, int res) {
  return actc == 0 ? pick_aux_res<0>(res) : actc == 1 ? pick_aux_res<1>(res) : pick_aux_res<2>(res);
}
int smem_for(int bk) { return bk == 64 ? KCfg<64>::kSmemBytes : bk == 32 ? KCfg<32>::kSmemBytes : KCfg<16>::kSmemBytes; }

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)ptr;
  }
  return fn;
}

// NHWC fp16 activation slice -> 4-D map {C, W, H, N}; strides in elements of the *buffer*.
CUtensorMapSwizzle swizzle_for(int box_c) {
  return box_c == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : box_c == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

int encode_act_map(CUtensorMap* m, const void* base, int C, int W, int H, int N, long long sW, long long sH, long long sN,
                   int box_c, int box_w, int box_h) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return his_set_error(HIS_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)sW * 2, (cuuint64_t)sH * 2, (cuuint64_t)sN * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  if (((uintptr_t)base & 15) || (strides[0] & 15) || (strides[1] & 15) || (strides[2] & 15))
    return his_set_error(HIS_ERR_INVALID_ARG, "activation slice is not 16-byte aligned (channel offset/stride must be multiples of 8)");
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(box_c), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[128]; snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(act) failed: %d", (int)r);
    return his_set_error(HIS_ERR_DRIVER, buf);
  }
  return HIS_OK;
}

int encode_weight_map(CUtensorMap* m, const void* base, int K, long long rows, int box_rows, int bk) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return his_set_error(HIS_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  if (((uintptr_t)base & 15) || (strides[0] & 15)) return his_set_error(HIS_ERR_INVALID_ARG, "packed weights must be 16-byte aligned, K % 8 == 0");
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(bk), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[128]; snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
    return his_set_error(HIS_ERR_DRIVER, buf);
  }
  return HIS_OK;
}

struct ConvGemmPlan {
  CUtensorMap tmA, tmB, tmO[4], tmR;
  ConvGemmParams p;
  int grid, bk, smem, actc, res_mode, transposed, cin_pad;
  ConvGemmKernel kernel;
};

// K-block width: the largest of {64,32,16} whose padded K stays within 25 % of the best padding
int pick_bk(int cin) {
  const int c16 = (cin + 15) / 16 * 16, c32 = (cin + 31) / 32 * 32, c64 = (cin + 63) / 64 * 64;
  if (c64 * 4 <= c16 * 5) return 64;
  if (c32 * 4 <= c16 * 5) return 32;
  return 16;
}

int g_num_sms = 0;

}  // namespace

extern "C" {

// See include/his_b200.h for the contract.
int his_conv_gemm_tile_n(int cout, int* n_tiles, int* block_n) {
  if (cout <= 0) return HIS_ERR_INVALID_ARG;
  int c16 = (cout + 15) / 16 * 16;
  if (c16 <= 256) { *n_tiles = 1; *block_n = c16; return HIS_OK; }
  int nt = (c16 + 255) / 256;
  int bn = ((c16 + nt - 1) / nt + 31) / 32 * 32;   // multi-tile: block_n % 32 == 0 so 32-wide store chunks never overlap
  *n_tiles = (c16 + bn - 1) / bn; *block_n = bn;
  return HIS_OK;
}

int his_conv_gemm_create(void** out_plan,
                         const void* in, int n_img, int H, int W, int cin, int in_cs,
                         const void* w_packed, int cin_pad,
                         void* out, int cout, int out_cs,
                         const void* res, int res_cs,
                         const float* shift,
                         int ksize, int transposed, int act, float act_beta, int res_mode) {
  if (!out_plan || !in || !w_packed || !out || !shift) return his_set_error(HIS_ERR_INVALID_ARG, "null pointer");
  if (!(ksize == 1 || ksize == 3) || (transposed && ksize != 1)) return his_set_error(HIS_ERR_UNSUPPORTED, "ksize must be 1 or 3 (transposed: k2s2 packed as 4 1x1 groups)");
  if (res_mode != HIS_RES_NONE && !res) return his_set_error(HIS_ERR_INVALID_ARG, "res_mode set without residual tensor");
  if (res_mode != HIS_RES_NONE && transposed) return his_set_error(HIS_ERR_UNSUPPORTED, "residual with transposed conv");
  if ((in_cs % 8) || (out_cs % 8) || (cin_pad % 8) || cin_pad < cin) return his_set_error(HIS_ERR_INVALID_ARG, "channel strides must be multiples of 8");
  if (g_num_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return his_set_error(HIS_ERR_NO_DEVICE, "no CUDA device");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return his_set_error(HIS_ERR_NO_DEVICE, "no CUDA device");
    if (prop.major != 10) return his_set_error(HIS_ERR_UNSUPPORTED, "conv_gemm_sm100 needs a compute-capability 10.x device (B200)");
    for (int bk = 16; bk <= 64; bk *= 2)
      for (int a = 0; a < 3; ++a)
        for (int r = 0; r < 3; ++r)
          if (cudaFuncSetAttribute(pick_kernel(bk, a, r), cudaFuncAttributeMaxDynamicSharedMemorySize, smem_for(bk)) != cudaSuccess)
            return his_set_error(HIS_ERR_LAUNCH, "cannot raise dynamic shared memory limit");
    for (int bk = 16; bk <= 64; bk *= 2)
      for (int r = 0; r < 2; ++r)
        if (cudaFuncSetAttribute(pick_tail_kernel(bk, r), cudaFuncAttributeMaxDynamicSharedMemorySize, smem_for(bk)) != cudaSuccess)
          return his_set_error(HIS_ERR_LAUNCH, "cannot raise dynamic shared memory limit");
    for (int a = 0; a < 3; ++a)
      for (int r = 0; r < 3; ++r)
        if (cudaFuncSetAttribute(pick_aux_kernel(a, r), cudaFuncAttributeMaxDynamicSharedMemorySize, smem_for(64)) != cudaSuccess)
          return his_set_error(HIS_ERR_LAUNCH, "cannot raise dynamic shared memory limit");
    g_num_sms = prop.multiProcessorCount;
  }
  ConvGemmPlan* pl = new ConvGemmPlan();
  memset(pl, 0, sizeof(*pl));
  ConvGemmParams& p = pl->p;
  p.n_img = n_img; p.H = H; p.W = W; p.ksize = ksize;
  // pick the 128-pixel rectangle with the least padded area
  long long best = -1;
  for (int bw = 1; bw <= 128; bw <<= 1) {
    int bh = 128 / bw;
    long long area = (long long)his_div_up(W, bw) * bw * his_div_up(H, bh) * bh;
    if (best < 0 || area < best || (area == best && bw >= 8 && bw <= 32)) { best = area; p.bw = bw; p.bh = bh; }
  }
  p.tiles_x = his_div_up(W, p.bw); p.tiles_y = his_div_up(H, p.bh);
  const int bk = pick_bk(cin);
  pl->bk = bk;
  p.kblocks_per_tap = his_div_up(cin, bk);
  his_conv_gemm_tile_n(cout, &p.n_tiles, &p.block_n);
  p.groups = transposed ? 4 : 1;
  p.cout_slab = p.n_tiles * p.block_n;
  p.num_work = n_img * p.tiles_y * p.tiles_x * p.n_tiles * p.groups;
  int actc = ACTC_CLAMP;
  p.act_lo = -INFINITY; p.act_beta = 1.0f; p.act_mul_x = 0;
  switch (act) {
    case HIS_ACT_NONE: break;
    case HIS_ACT_RELU: p.act_lo = 0.0f; break;
    case HIS_ACT_SILU: actc = ACTC_SIGMOID; p.act_mul_x = 1; break;
    case HIS_ACT_SIGMOID: actc = ACTC_SIGMOID; break;
    case HIS_ACT_SWISH: actc = ACTC_SIGMOID; p.act_mul_x = 1; p.act_beta = act_beta; break;
    case HIS_ACT_GELU: actc = ACTC_GELU; break;
    default: delete pl; return his_set_error(HIS_ERR_INVALID_ARG, "unknown activation code");
  }
  if (res_mode < 0 || res_mode > 2) { delete pl; return his_set_error(HIS_ERR_INVALID_ARG, "unknown res_mode"); }
  pl->kernel = pick_kernel(bk, actc, res_mode);
  pl->smem = smem_for(bk);
  p.stage_bytes = (kBlockM * bk * 2 + p.block_n * bk * 2 + 1023) / 1024 * 1024;
  p.stages = KCfg<64>::kRingBytes / p.stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  if (const char* e = getenv("HIS_GEMM_STAGES")) { int v = atoi(e); if (v >= 2 && v < p.stages) p.stages = v; }
  pl->actc = actc; pl->res_mode = res_mode; pl->transposed = transposed; pl->cin_pad = cin_pad;
  p.tail_c = 0; p.store_main = 1; p.aux_out = nullptr; p.cout = cout;
  p.shift = shift;
  int taps = ksize * ksize;
  int rc;
  if ((rc = encode_act_map(&pl->tmA, in, cin, W, H, n_img, in_cs, (long long)W * in_cs, (long long)H * W * in_cs, bk, p.bw, p.bh))) { delete pl; return rc; }
  if ((rc = encode_weight_map(&pl->tmB, w_packed, cin_pad, (long long)p.groups * taps * p.cout_slab, p.block_n, bk))) { delete pl; return rc; }
  if (!transposed) {
    if ((rc = encode_act_map(&pl->tmO[0], out, cout, W, H, n_img, out_cs, (long long)W * out_cs, (long long)H * W * out_cs, kChunkC, p.bw, p.bh))) { delete pl; return rc; }
    pl->tmO[1] = pl->tmO[2] = pl->tmO[3] = pl->tmO[0];
  } else {
    for (int g = 0; g < 4; ++g) {
      int dy = g >> 1, dx = g & 1;
      const __half* base = (const __half*)out + ((long long)dy * 2 * W + dx) * out_cs;
      if ((rc = encode_act_map(&pl->tmO[g], base, cout, W, H, n_img, 2LL * out_cs, 4LL * W * out_cs, 4LL * H * W * out_cs, kChunkC, p.bw, p.bh))) { delete pl; return rc; }
    }
  }
  if (res_mode != HIS_RES_NONE) {
    if ((rc = encode_act_map(&pl->tmR, res, cout, W, H, n_img, res_cs, (long long)W * res_cs, (long long)H * W * res_cs, kChunkC, p.bw, p.bh))) { delete pl; return rc; }
  } else {
    pl->tmR = pl->tmO[0];
  }
  pl->grid = p.num_work < g_num_sms ? p.num_work : g_num_sms;
  *out_plan = pl;
  return HIS_OK;
}

int his_conv_gemm_set_tail(void* plan, const float* tail_w, float tail_b0, float tail_b1, int tail_c, int tail_sigmoid, float* tail_out,
                           int store_main) {
  if (!plan || !tail_w || !tail_out) return his_set_error(HIS_ERR_INVALID_ARG, "set_tail: null pointer");
  ConvGemmPlan* pl = (ConvGemmPlan*)plan;
  if (tail_c < 1 || tail_c > 2) return his_set_error(HIS_ERR_INVALID_ARG, "set_tail: tail_c must be 1 or 2");
  if (pl->p.n_tiles != 1 || pl->transposed || pl->actc != ACTC_CLAMP || pl->res_mode == HIS_RES_MUL)
    return his_set_error(HIS_ERR_UNSUPPORTED, "set_tail: needs a single N tile, none/relu activation, no MUL operand, not transposed");
  pl->p.tail_w = tail_w; pl->p.tail_b0 = tail_b0; pl->p.tail_b1 = tail_b1; pl->p.tail_c = tail_c; pl->p.tail_sigmoid = tail_sigmoid;
  pl->p.tail_out = tail_out; pl->p.store_main = store_main;
  pl->kernel = pick_tail_kernel(pl->bk, pl->res_mode);
  return HIS_OK;
}

int his_conv_gemm_set_image_weights(void* plan, const void* w_packed_per_image) {
  if (!plan || !w_packed_per_image) return his_set_error(HIS_ERR_INVALID_ARG, "set_image_weights: null pointer");
  ConvGemmPlan* pl = (ConvGemmPlan*)plan;
  ConvGemmParams& p = pl->p;
  const long long rows = (long long)p.groups * p.ksize * p.ksize * p.cout_slab;
  if (rows * p.n_img >= (1LL << 31)) return his_set_error(HIS_ERR_UNSUPPORTED, "set_image_weights: too many weight rows");
  int rc = encode_weight_map(&pl->tmB, w_packed_per_image, pl->cin_pad, rows * p.n_img, p.block_n, pl->bk);
  if (rc) return rc;
  p.b_img_rows = (int)rows;
  return HIS_OK;
}

int his_conv_gemm_set_aux(void* plan, float* aux_out) {
  if (!plan || !aux_out) return his_set_error(HIS_ERR_INVALID_ARG, "set_aux: null pointer");
  ConvGemmPlan* pl = (ConvGemmPlan*)plan;
  if (pl->bk != 64 || pl->transposed || pl->p.tail_c)
    return his_set_error(HIS_ERR_UNSUPPORTED, "set_aux: needs a 64-wide K block (Cin >= 64), not transposed, no fused tail");
  pl->p.aux_out = aux_out;
  pl->kernel = pick_aux_kernel(pl->actc, pl->res_mode);
  return HIS_OK;
}

int his_conv_gemm_run(void* plan, void* stream) {
  if (!plan) return his_set_error(HIS_ERR_INVALID_ARG, "null plan");
  ConvGemmPlan* pl = (ConvGemmPlan*)plan;
  if (pl->p.num_work == 0) return HIS_OK;
  pl->kernel<<<pl->grid, kThreadsGemm, pl->smem, (cudaStream_t)stream>>>(pl->tmA, pl->tmB, pl->tmO[0], pl->tmO[1], pl->tmO[2], pl->tmO[3], pl->tmR,
                                                               pl->p);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_conv_gemm_destroy(void* plan) {
  delete (ConvGemmPlan*)plan;
  return HIS_OK;
}

// Algorithmic FLOPs of one run (2*MAC over true channel counts are computed by the caller); this
// returns the padded MMA work actually issued, for roofline bookkeeping.
long long his_conv_gemm_issued_macs(void* plan) {
  ConvGemmPlan* pl = (ConvGemmPlan*)plan;
  const ConvGemmParams& p = pl->p;
  return (long long)p.num_work * kBlockM * p.block_n * (long long)(p.ksize * p.ksize * p.kblocks_per_tap * pl->bk);
}

}  // extern "C"
