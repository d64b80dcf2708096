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
//    accumulator), so two tiles drain concurrently while the MMAs of later tiles run.  (Narrow layers and the 1x1 layers run three
//    groups.)  A full chunk takes a straight-line body (ld.shared / math / st.shared per 16-channel half); the residual box of
//    chunk j+1 is requested from the middle of chunk j; the issuer drains earlier stores before the barrier of its own store, so
//    a chunk starts without a wait -- see the comments in the epilogue and DESIGN.md for the clock-stamp measurements behind this.
//  * conv_transpose k2s2 = 4 independent 1x1 GEMMs ("groups"), each scattered through its own
//    strided output tensor map (pixel (2y+dy, 2x+dx)).
//
// Warp roles (384 threads, 1 CTA/SM, persistent over work items):
//   warp 0: TMA producer   warp 1: MMA issuer   warp 2: TMEM allocator   warps 4-7 / 8-11: epilogue groups 0 / 1.
#pragma once
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
constexpr int kMaxEpiGroups = 3;                       // epilogue warpgroups: 2 (384 threads) or 3 (512 threads), see ConvGemmParams::epi_groups
constexpr int kShiftBytes = (1 + kMaxEpiGroups + 2) * 256 * 4;   // per-channel shift of the (single) N tile, one residual-scale row per epilogue
                                                                 // group, two weight rows of a fused 1x1 tail
constexpr int kThreadsGemm = 128 + 128 * kMaxEpiGroups;      // launch bound; a plan launches 128 + 128 * epi_groups threads
constexpr int kTmemCols = 512;
constexpr int kMaxStages = 48;
constexpr int kSmemBudget = 227 * 1024;
constexpr int kBarrierBytes = 1024;
// Halo mode (3x3 convs): the A operand of a tile is the (bh+2) x (bw+2) input window, loaded ONCE per K block by two
// cp.async producer warps into an un-swizzled K-major layout [8-channel plane][halo row][halo col][16 B]; the nine filter taps
// are nine start addresses into that window (8-row core-matrix groups = tile rows, SBO = one halo row).  The 9x re-read of A
// through L2 -- the bound of every layer with Cout < ~192 (L2 serves ~42 B/clk/SM) -- disappears.
constexpr int kHaloBw = 8, kHaloBh = 16;               // tile = 8 x 16 output pixels (row m = y*8 + x)
constexpr int kHaloW = kHaloBw + 2, kHaloH = kHaloBh + 2, kHaloPix = kHaloW * kHaloH;
constexpr int kPlaneBytes = (kHaloPix + 1) * 16;       // 2896: +16 B so the planes of one pixel fall in distinct 16-byte bank groups
constexpr int kAProducerThreads = 64;                  // warps 2 and 3
constexpr int kAuxStagingBytes = kChunkC * kBlockM * 4; // EPI_AUX: one [32 channels][128 pixels] fp32 tile per epilogue group
constexpr int kMaxAStages = 24, kMaxBStages = 16;

// K-block width BK (fp16 elements) selects the shared-memory swizzle: one row of the operand tile is BK*2 bytes.
template <int BK> struct KCfg {
  static constexpr int kABytes = kBlockM * BK * 2;
  static constexpr int kBBytesMax = 256 * BK * 2;
  // the ring gets whatever the 227 KB leave after the staging buffers, barriers and the 1 KB alignment slack; its depth is a
  // run-time parameter (stage = A + the layer's actual B tile), so narrow layers keep many more loads in flight
  static constexpr int kRingBytes = kSmemBudget - 4 * kStagingBytes - kShiftBytes - 1024 - kBarrierBytes;      // two epilogue groups
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
                             // (halo mode: the B ring; stage = taps_per_b weight tiles)
  // halo mode
  const __half* in; long long in_sn; int in_cs, cin;   // raw NHWC input (element strides) for the cp.async producers
  int a_stages, a_stage_bytes, taps_per_b;
  // direct (register -> global) epilogue path for clipped chunks / tiles (non-transposed layers)
  __half* out; const __half* res; int out_cs, res_cs, direct_ok;
  int n_acc, acc_stride; // TMEM accumulator ring: n_acc (even, <= 8) buffers of acc_stride columns
  int taps_per_box;     // halo mode: taps per weight TMA box (taps_per_b / taps_per_box boxes fill one B stage)
  int b_resident;       // halo mode: the whole weight set stays in shared memory (loaded once per CTA)
  float inv_tiles_x, inv_tiles_y;   // fast work decode (num_work < 2^21, single N tile, no groups)
  int fast_decode;
  // per-pixel (GEMM row) extras: y = act(row_scale[pix]*acc + shift ...) and per-pixel channel mean / max of y -> stats[pix][2]
  // (SpatialAttentionModule, attention_modules.py:67-113: the gate multiplies the next conv's input, the statistics feed it)
  const float* row_scale; float* stats_out;
  const float* res_scale;   // [n_img][cout] multiplier of the residual / multiplicand operand (ChannelAttention gate folded into the block)
  // halo mode, fused nearest 2x upsample (smp UnetDecoderBlock: F.interpolate(x, nearest) then cat with the skip): channels
  // [0, up_split) are gathered from the low-resolution tensor `up_in` [n, H/2, W/2, up_cs] at (y>>1, x>>1), the rest from `in`
  const __half* up_in; long long up_sn; int up_cs, up_split;
  int pair;             // 1: CTA pairs (cluster of 2) run cta_group::2 MMAs, M = 256 pixels x block_n, each CTA stages half of the weight tile
  int debug;            // HIS_GEMM_DEBUG bit mask (tuning experiments only): 1 no epilogue stores, 2 no A loads, 4 no B loads, 8 no MMAs,
                        // 16 epilogue time stamps, 32 general chunk body everywhere, 64 residual chunks loaded at the top of their own chunk
                        // + wait / barrier at the top of every chunk, 128 general body for tiles with a residual scale
  // activation, compile-time class + runtime parameters:
  //   CLAMP:   y = max(y, act_lo)                  (none: -inf, relu: 0)
  //   SIGMOID: s = 1/(1+exp(-act_beta*y)); y = act_mul_x ? y*s : s   (sigmoid / silu / swish(beta))
  //   GELU:    exact erf form
  float act_lo, act_beta, act_nb2;
  int act_mul_x;
  const float* shift;   // [cout_slab]  (conv bias + folded-BatchNorm shift; the BN scale lives in the weights)
  // fused 1x1 tail to <= 2 channels (TAIL kernels): tail[o] = sum_c y[c]*tail_w[o][c] + tail_b[o], NCHW fp32 out
  const float* tail_w;  // [tail_c][cout_slab]
  float* tail_out;      // [n_img][tail_c][H][W]
  float tail_b0, tail_b1;
  int tail_c, tail_sigmoid, store_main;
  // fp32 NCHW copy of the layer output (EPI_AUX kernels): aux[n][c][y][x] = y, or the gate itself (before the product) for RES_MUL
  float* aux_out;
  int aux_bufs;         // export staging tiles per epilogue group (2: the TMA store of chunk j drains while chunk j+1 is written)
  int aux_tma;          // 1: the export leaves through a [32 ch][128 px] fp32 staging tile and ONE TMA store per chunk (coalesced
                        // 64-byte rows per channel) instead of 32 strided 4-byte stores per thread; needs W % 4 == 0
  int cout;             // true output channel count
  // split-fp16 ("strict" precision) operands, SPLIT kernels only: every activation is a pair of fp16 planes x = hi + lo that live
  // lo elements apart inside one pixel ([hi channels | lo channels], lo = pixel stride / 2), the packed weights are [W_hi | W_lo]
  // along K (cin_pad1 elements each).  Per tap the K loop runs 3 * nblk_phys blocks: A_hi.W_hi, A_lo.W_hi, A_hi.W_lo (the
  // lo.lo term is below 2^-22 relative); products of fp16 pairs are exact in the fp32 accumulator, so the result carries ~21
  // significant bits.  The epilogue emits hi = fp16(y), lo = fp16(y - hi).  kblocks_per_tap holds the VIRTUAL count (3x).
  int nblk_phys, cin_pad1, in_lo, up_lo, out_lo, res_lo;
  // conv_transpose k2s2 with merged phases: one N tile holds phase_merge (2 or 4) of the four (dy,dx) phase GEMMs, phase_slab
  // columns each, so the A tile is read 4 / phase_merge times instead of four and the MMAs run at N = 256; every 32-channel
  // epilogue chunk goes to its phase's strided output map.  phase_merge == 1: one group per phase (or not transposed).
  int phase_merge, phase_slab;
  // epilogue warpgroups (2 or 3), each with two staging buffers; tiles go to the groups round robin.  Narrow layers are bound by
  // the per-tile latency of one warpgroup's dependent instruction chain (wait, tcgen05.ld, math, stores: ~1 900 clk per tile
  // and group measured with every memory operation removed), not by any throughput: a third group gives 1.5x there.
  int epi_groups;
  // halo mode: 1 = the (bh+2) x (bw+2) input window of a K block arrives as ONE 5-D TMA box {8 ch, 10 cols, 18 rows, BK/8 planes, 1}
  // of the map {8, W, H, C/8, N} (out-of-image coordinates and the channel tail zero-filled by the TMA unit), issued by one
  // lane of warp 2; 0 = the two cp.async producer warps (needed for the fused nearest upsample and the split-fp16 planes).
  // plane_bytes = pitch of an 8-channel plane in the window: 180*16 for the TMA box, +16 for the cp.async writes' bank spread.
  int a_tma, plane_bytes;
  // LayerNorm2d statistics from the epilogue (hed/model.py:18-38: mean / variance over (C,H,W) per sample): every work item
  // writes (sum, sum of squares) of the values it stores -- the fp16-rounded ones, i.e. what the normalise pass will read -- to
  // ln_partials[work item][2] (double); a sample's work items are contiguous, his_layernorm2d_act adds them in order.
  double* ln_partials;
  // HIS_GEMM_DEBUG & 16 (tuning only): thread 128 of CTA 0 writes clock64() stamps of its first 64 tiles here, 32 slots per tile
  unsigned long long* dbg_ts;
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
// 5-D forms of the SPLIT kernels: activation maps are {C, part (hi / lo), W, H, N}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
// ---- CTA-pair (cta_group::2) forms: TMA loads of both CTAs complete on the LEADER's barrier, the leader's MMA thread issues one
// M = 256 instruction that reads A (own 128 rows) and half of B from each CTA's shared memory and accumulates into each CTA's TMEM;
// its commit arrives on the same-offset barrier of both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t cluster_bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(cluster_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t cluster_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(cluster_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}

__device__ __forceinline__ void umma_f16_acc_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc)
      : "memory");
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

// K-major un-swizzled (INTERLEAVE) descriptor of the halo window: core matrix = 8 rows x 16 B, rows 16 B apart;
// SBO = bytes between 8-row groups (one halo row), LBO = bytes between the two 8-channel planes of one K=16 step.
__device__ __forceinline__ uint64_t make_halo_desc(uint32_t saddr, int plane_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(plane_bytes >> 4) << 16;
  d |= (uint64_t)((kHaloW * 16) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// kind::f16 instruction descriptor: D=f32, A=B=f16, both K-major, M=128, N=n.
__device__ __forceinline__ uint32_t make_idesc_f16(int n, int m = kBlockM) {
  uint32_t d = 0;
  d |= 1u << 4;                       // c_format = F32
  d |= 0u << 7;                       // a_format = F16
  d |= 0u << 10;                      // b_format = F16
  d |= (uint32_t)(n >> 3) << 17;      // n_dim
  d |= (uint32_t)(m >> 4) << 24;        // m_dim (256 = CTA pair)
  return d;
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// accumulate form (enable-input-d = true) without the predicate set-up: all but the first MMA of a tile
__device__ __forceinline__ void umma_f16_acc(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

template <bool P>
__device__ __forceinline__ void umma_x(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (P) umma_f16_2sm(tmem_d, adesc, bdesc, idesc, accumulate); else umma_f16(tmem_d, adesc, bdesc, idesc, accumulate);
}
template <bool P>
__device__ __forceinline__ void umma_acc_x(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  if (P) umma_f16_acc_2sm(tmem_d, adesc, bdesc, idesc); else umma_f16_acc(tmem_d, adesc, bdesc, idesc);
}
template <bool P>
__device__ __forceinline__ void umma_commit_x(uint32_t bar) {
  if (P) umma_commit_2sm(bar); else umma_commit(bar);
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

// explicit shared-window 16-byte accesses (the epilogue's pointers are otherwise generic: LD / ST instead of LDS / STS)
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts32f(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int ACTC>
__device__ __forceinline__ float epi_act(float y, const ConvGemmParams& p) {
  if (ACTC == ACTC_CLAMP) return fmaxf(y, p.act_lo);
  if (ACTC == ACTC_SIGMOID) {
    const float s = his_sigmoid_fast(y, p.act_nb2);       // act_nb2 = -act_beta * log2(e)
    return p.act_mul_x ? y * s : s;
  }
  return 0.5f * y * (1.0f + erff(y * 0.70710678118654752f));
}

struct WorkItem { int img, y0, x0, n_tile, group; };

__device__ __forceinline__ WorkItem decode_work(const ConvGemmParams& p, int w) {
  WorkItem it;
  if (p.fast_decode) {   // exact for w < 2^21: |fl((w+.5)*inv) - (w+.5)/d| < .5/d
    const int r = __float2int_rz(((float)w + 0.5f) * p.inv_tiles_x);      // w / tiles_x
    const int tx = w - r * p.tiles_x;
    const int im = __float2int_rz(((float)r + 0.5f) * p.inv_tiles_y);     // r / tiles_y
    it.n_tile = 0; it.group = 0; it.img = im; it.y0 = (r - im * p.tiles_y) * p.bh; it.x0 = tx * p.bw;
    return it;
  }
  it.n_tile = w % p.n_tiles; w /= p.n_tiles;
  it.group = w % p.groups;   w /= p.groups;
  int tx = w % p.tiles_x;    w /= p.tiles_x;
  int ty = w % p.tiles_y;    w /= p.tiles_y;
  it.img = w; it.y0 = ty * p.bh; it.x0 = tx * p.bw;
  return it;
}

// ------------------------------------------------------------------------------------ kernel
enum { EPI_PLAIN = 0, EPI_TAIL = 1, EPI_AUX = 2 };

template <int BK, int ACTC, int RES, int EPI, bool HALO, bool PAIR = false, bool SPLIT = false>
__global__ void __launch_bounds__(kThreadsGemm, 1)
conv_gemm_sm100_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1,
                       const __grid_constant__ CUtensorMap tmO2, const __grid_constant__ CUtensorMap tmO3,
                       const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmAux, const ConvGemmParams p) {
  using Cfg = KCfg<BK>;
  const int kStages = p.stages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_base = smem_base;
  const uint32_t staging_base = smem_base + kStages * p.stage_bytes;
  const uint32_t shift_base = staging_base + (uint32_t)(2 * p.epi_groups) * kStagingBytes;
  const uint32_t bar_base = shift_base + kShiftBytes;
  // barrier slots (8 B each): full[kStages] empty[kStages] tmem_full[2] tmem_empty[2] res_full[4], the TMEM pointer, then (halo
  // mode) afull[a_stages] aempty[a_stages]; the halo ring itself follows the barrier block
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 8 + a); };
  auto res_bar = [&](int b) { return bar_base + 8u * (2 * kStages + 16 + b); };          // 2 per epilogue group (<= 6)
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 22);
  auto afull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 23 + s); };
  auto aempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 23 + p.a_stages + s); };
  auto apeer_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 23 + 2 * p.a_stages + s); };   // pair: the peer's window is complete
  const int n_acc = p.n_acc;
  const uint32_t a_base = bar_base + kBarrierBytes;
  // EPI_AUX: fp32 export staging (one tile per epilogue group) behind the halo ring
  const uint32_t aux_base = a_base + (uint32_t)(HALO ? p.a_stages * p.a_stage_bytes : 0);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));   // generic pointer to the aligned base

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int taps = p.ksize * p.ksize;
  const int kiters = taps * p.kblocks_per_tap;
  // PAIR kernels hold cta_group::2 instructions and can only be launched as clusters of two CTAs
  constexpr bool pair = PAIR;
  const uint32_t cta_rank = pair ? cluster_ctarank() : 0u;
  // bytes of the weight tile this CTA stages (a pair member holds half of the N rows)
  const uint32_t b_bytes = (uint32_t)(pair ? p.block_n >> 1 : p.block_n) * BK * 2;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB); prefetch_tmap(&tmO0);
    if (RES) prefetch_tmap(&tmR);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 8; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), pair ? 8 : 4); }   // pair: both CTAs' epilogue warps
    for (int a = 0; a < 2 * kMaxEpiGroups; ++a) mbar_init(res_bar(a), 1);
    if (HALO)
      for (int s = 0; s < p.a_stages; ++s) { mbar_init(afull_bar(s), p.a_tma ? 1 : kAProducerThreads); mbar_init(aempty_bar(s), 1); mbar_init(apeer_bar(s), 1); }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (pair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  // single N tile: the per-channel shift stays in shared memory for the whole kernel
  float* s_shift = reinterpret_cast<float*>(smem_gen + (shift_base - smem_base));
  if (p.n_tiles == 1)
    for (int i = threadIdx.x; i < p.block_n; i += blockDim.x) s_shift[i] = __ldg(p.shift + (p.phase_merge > 1 ? i % p.phase_slab : i));
  // fused 1x1 tail: its weight rows live in rows 4 / 5 of the shift block
  constexpr uint32_t kTailRow = (1 + kMaxEpiGroups) * 1024;       // byte offset of tail row 0 from the shift row
  const bool tail_smem = EPI == EPI_TAIL && p.n_tiles == 1;
  if (tail_smem)
    for (int i = threadIdx.x; i < p.tail_c * p.cout_slab; i += blockDim.x)
      s_shift[(1 + kMaxEpiGroups) * 256 + (i / p.cout_slab) * 256 + (i % p.cout_slab)] = __ldg(p.tail_w + i);
  tc_fence_before();
  __syncthreads();
  if (pair) cluster_sync_all();        // the peer's barriers are initialised and its TMEM is allocated before anything remote happens
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  if (HALO) {
    const int nblk = p.kblocks_per_tap;                  // K blocks of BK channels
    const int tgroups = 9 / p.taps_per_b;                // B stages per K block
    if (warp == 0) {
      // ================================ halo mode: weight-tile TMA producer ================================
      int stage = 0; uint32_t phase = 0;
      const uint32_t stage_tx = (uint32_t)p.taps_per_b * b_bytes;
      // K coordinate of (virtual) K block cb in the packed weights; SPLIT: blocks [2*nblk_phys, 3*nblk_phys) read W_lo
      auto bk_coord = [&](int cb) {
        if (!SPLIT) return cb * BK;
        const int s3 = cb / p.nblk_phys, cbp = cb - s3 * p.nblk_phys;
        return cbp * BK + (s3 == 2 ? p.cin_pad1 : 0);
      };
      if (p.b_resident) {       // every (K block, tap) weight tile once, all on full_bar(0)
        if (elect_one()) {
          mbar_expect_tx(full_bar(0), (uint32_t)(nblk * 9) * b_bytes);
          for (int cb = 0; cb < nblk; ++cb)
            for (int tap = 0; tap < 9; tap += p.taps_per_box)
              tma_load_2d(stage_base + (uint32_t)(cb * 9 + tap) * b_bytes, &tmB, full_bar(0), bk_coord(cb), tap * p.cout_slab);
        }
        __syncwarp();
      }
      for (int w = blockIdx.x; w < p.num_work && !p.b_resident; w += gridDim.x) {
        const WorkItem it = decode_work(p, w);
        const int brow0 = it.n_tile * p.block_n + it.img * p.b_img_rows;
        for (int cb = 0; cb < nblk; ++cb)
          for (int tg = 0; tg < tgroups; ++tg) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            if (pair) {
              if (elect_one()) {       // both CTAs stage their half of every tap's weight rows; completion on the leader's barrier
                const uint32_t sb = stage_base + stage * p.stage_bytes;
                const uint32_t lead_full = mapa_shared(full_bar(stage), 0);
                if (cta_rank == 0) mbar_expect_tx(full_bar(stage), 2u * stage_tx);
                for (int t = 0; t < p.taps_per_b; ++t)
                  tma_load_2d_2sm(sb + (uint32_t)t * b_bytes, &tmB, lead_full, bk_coord(cb),
                                  brow0 + (tg * p.taps_per_b + t) * p.cout_slab + (int)cta_rank * (p.block_n >> 1));
              }
            } else if (elect_one()) {
              const uint32_t sb = stage_base + stage * p.stage_bytes;
              if (p.debug & 4) mbar_arrive(full_bar(stage));
              else {
                mbar_expect_tx(full_bar(stage), stage_tx);
                // a box spans taps_per_box taps (> 1 only with a single N tile, where the taps' weight rows are contiguous)
                for (int t = 0; t < p.taps_per_b; t += p.taps_per_box)
                  tma_load_2d(sb + (uint32_t)t * b_bytes, &tmB, full_bar(stage), bk_coord(cb), brow0 + (tg * p.taps_per_b + t) * p.cout_slab);
              }
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
      }
    } else if (warp == 1) {
      // ================================ halo mode: MMA issuer ================================
      // One thread feeds the tensor pipe, so the loop is kept short: descriptors are a constant high word plus a 14-bit
      // address that advances by adds, the K=16 steps of a block are unrolled, and a B stage carries up to nine taps.
      const uint32_t idesc = make_idesc_f16(p.block_n, pair ? 2 * kBlockM : kBlockM);
      const uint64_t adesc_hi = make_halo_desc(0, p.plane_bytes), bdesc_hi = make_kmajor_desc<BK>(0);
      const uint64_t k16_units = (uint64_t)(2 * (p.plane_bytes >> 4));       // two 8-channel planes per K = 16 step
      if (pair && cta_rank != 0) {
        // peer of a pair: this warp only relays "my window of stage s is complete" to the leader's MMA thread
        int astage = 0; uint32_t aphase = 0;
        for (int w = blockIdx.x; w < p.num_work; w += gridDim.x)
          for (int cb = 0; cb < nblk; ++cb) {
            mbar_wait(afull_bar(astage), aphase);
            fence_proxy_async();
            if (elect_one()) mbar_arrive_cluster(mapa_shared(apeer_bar(astage), 0));
            __syncwarp();
            if (++astage == p.a_stages) { astage = 0; aphase ^= 1; }
          }
      }
      const int T = p.taps_per_b, cin = p.cin, a_stages = p.a_stages;
      const uint32_t b_tap_units = b_bytes >> 4, b_stage_units = (uint32_t)p.stage_bytes >> 4, a_stage_units = (uint32_t)p.a_stage_bytes >> 4;
      const uint32_t a_units0 = a_base >> 4, b_units0 = stage_base >> 4;
      int stage = 0, astage = 0; uint32_t phase = 0, aphase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      uint32_t a_units = a_units0, b_units = b_units0;
      if (p.b_resident) { mbar_wait(full_bar(0), 0); tc_fence_after(); }      // weights: loaded once, never released
      for (int w = blockIdx.x; w < p.num_work && cta_rank == 0; w += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_stride);
        for (int cb = 0; cb < nblk; ++cb) {
          mbar_wait(afull_bar(astage), aphase);
          if (pair) mbar_wait(apeer_bar(astage), aphase);
          fence_proxy_async();                 // cp.async (generic proxy) writes -> tcgen05 (async proxy) reads
          tc_fence_after();
          const int nk16 = min(BK / 16, (cin - (SPLIT ? cb % p.nblk_phys : cb) * BK + 15) >> 4);
          if (p.b_resident) b_units = b_units0 + (uint32_t)(cb * 9) * b_tap_units;
          for (int tg = 0; tg < tgroups; ++tg) {
            if (!p.b_resident) { mbar_wait(full_bar(stage), phase); tc_fence_after(); }
            if (T == 9) {
              // all nine taps in one stage (weight-resident / narrow layers): fully unrolled, descriptor offsets are constants
              if (elect_one()) {
                const uint64_t ad0 = adesc_hi | a_units, bd0 = bdesc_hi | b_units;
                if (!(p.debug & 8)) {
#pragma unroll
                  for (int t = 0; t < 9; ++t) {
                    const uint64_t ad = ad0 + (uint64_t)((t / 3) * kHaloW + (t % 3)), bd = bd0 + (uint64_t)t * b_tap_units;
                    if (t == 0) umma_x<pair>(d_tmem, ad, bd, idesc, cb ? 1u : 0u);
                    else umma_acc_x<pair>(d_tmem, ad, bd, idesc);
#pragma unroll
                    for (int kk = 1; kk < BK / 16; ++kk)
                      if (kk < nk16) umma_acc_x<pair>(d_tmem, ad + (uint64_t)kk * k16_units, bd + (uint64_t)(kk * 2), idesc);
                  }
                }
                if (!p.b_resident) umma_commit_x<pair>(empty_bar(stage));
                umma_commit_x<pair>(aempty_bar(astage));
                if (cb == nblk - 1) umma_commit_x<pair>(tfull_bar(acc));
              }
            } else if (elect_one()) {
              const int tap0 = tg * T;
              const int dy0 = (tap0 * 11) >> 5;                       // tap0 / 3 for 0..8
              uint32_t au = a_units + (uint32_t)(dy0 * kHaloW + (tap0 - dy0 * 3)), bu = b_units;
              int dx = tap0 - dy0 * 3;
              for (int t = 0; t < T; ++t) {
                const uint64_t ad = adesc_hi | au, bd = bdesc_hi | bu;
                if (!(p.debug & 8)) {
                  umma_x<pair>(d_tmem, ad, bd, idesc, (cb | tg | t) ? 1u : 0u);
#pragma unroll
                  for (int kk = 1; kk < BK / 16; ++kk)
                    if (kk < nk16) umma_acc_x<pair>(d_tmem, ad + (uint64_t)kk * k16_units, bd + (uint64_t)(kk * 2), idesc);
                }
                bu += b_tap_units; ++au;
                if (++dx == 3) { dx = 0; au += kHaloW - 3; }
              }
              if (!p.b_resident) umma_commit_x<pair>(empty_bar(stage));
              if (tg == tgroups - 1) {
                umma_commit_x<pair>(aempty_bar(astage));
                if (cb == nblk - 1) umma_commit_x<pair>(tfull_bar(acc));
              }
            }
            __syncwarp();
            if (!p.b_resident) {
              b_units += b_stage_units;
              if (++stage == kStages) { stage = 0; phase ^= 1; b_units = b_units0; }
            } else {
              b_units += (uint32_t)T * b_tap_units;
            }
          }
          a_units += a_stage_units;
          if (++astage == a_stages) { astage = 0; aphase ^= 1; a_units = a_units0; }
        }
        if (++acc == n_acc) { acc = 0; acc_phase ^= 1; }
      }
    } else if (warp < 4 && p.a_tma) {
      // ================================ halo mode: the input window as one 5-D TMA box per K block ================================
      if (warp == 2) {
        int astage = 0; uint32_t aphase = 0;
        const uint32_t tx = (uint32_t)(BK / 8) * (uint32_t)p.plane_bytes;
        for (int w = blockIdx.x; w < p.num_work; w += gridDim.x) {
          const WorkItem it = decode_work(p, w);
          for (int cb = 0; cb < nblk; ++cb) {
            mbar_wait(aempty_bar(astage), aphase ^ 1);
            if (elect_one()) {
              mbar_expect_tx(afull_bar(astage), tx);
              tma_load_5d(a_base + astage * p.a_stage_bytes, &tmA, afull_bar(astage), 0, it.x0 - 1, it.y0 - 1, cb * (BK / 8), it.img);
            }
            __syncwarp();
            if (++astage == p.a_stages) { astage = 0; aphase ^= 1; }
          }
        }
      }
    } else if (warp < 4) {
      // ================================ halo mode: cp.async producers of the input window ================================
      constexpr int CL = BK / 8;                           // chunk lanes: the 16-byte pieces of one pixel in this K block
      constexpr int PL = kAProducerThreads / CL;           // pixels in flight per pass
      constexpr int NIT = (kHaloPix + PL - 1) / PL;        // passes per stage (6 / 12 / 23)
      const int ptid = threadIdx.x - 64;
      const int c = ptid % CL, pl = ptid / CL;
      // per-thread tables, tile independent: window coordinates of the thread's pixels and their element offsets; the loop
      // below is fully unrolled so they live in registers and the passes are independent instructions streams
      int rel[NIT], hyx[NIT], rel_up[NIT];
      const int W2 = p.W >> 1;
#pragma unroll
      for (int k = 0; k < NIT; ++k) {
        const int px = pl + k * PL;
        const int hy = px / kHaloW, hx = px - hy * kHaloW;
        rel[k] = hy * p.W * p.in_cs + hx * p.in_cs;
        // low-resolution source: window origin (y0-1, x0-1) with y0, x0 even -> source pixel (y0/2 + ((hy-1)>>1), x0/2 + ((hx-1)>>1))
        rel_up[k] = ((hy - 1) >> 1) * W2 * p.up_cs + ((hx - 1) >> 1) * p.up_cs;
        hyx[k] = px < kHaloPix ? ((hy << 16) | hx) : (0x4000 << 16);     // beyond the window: never valid
      }
      int astage = 0; uint32_t aphase = 0;
      for (int w = blockIdx.x; w < p.num_work; w += gridDim.x) {
        const WorkItem it = decode_work(p, w);
        const int gy0 = it.y0 - 1, gx0 = it.x0 - 1;
        const __half* org = p.in + (long long)it.img * p.in_sn + ((long long)gy0 * p.W + gx0) * p.in_cs;   // window origin (may lie outside)
        const __half* org_up = p.up_in + (long long)it.img * p.up_sn + ((long long)(it.y0 >> 1) * W2 + (it.x0 >> 1)) * p.up_cs;
        for (int cb = 0; cb < nblk; ++cb) {
          mbar_wait(aempty_bar(astage), aphase ^ 1);
          // SPLIT: virtual block cb = part * nblk_phys + physical block; part 1 (A_lo) reads the lo plane of the same pixels
          const int s3 = SPLIT ? cb / p.nblk_phys : 0;
          const int ch = (SPLIT ? cb - s3 * p.nblk_phys : cb) * BK + c * 8;
          const bool chok = ch < p.cin && !(p.debug & 2);
          const uint32_t dst = a_base + astage * p.a_stage_bytes + (uint32_t)c * kPlaneBytes + (uint32_t)pl * 16u;
          const bool from_up = ch < p.up_split;                       // block uniform: up_split is a multiple of BK
          const __half* orgc = (from_up ? org_up : org) + ch + ((SPLIT && s3 == 1) ? (from_up ? p.up_lo : p.in_lo) : 0);
#pragma unroll
          for (int k = 0; k < NIT; ++k) {
            const bool ok = chok && (unsigned)(gy0 + (hyx[k] >> 16)) < (unsigned)p.H && (unsigned)(gx0 + (hyx[k] & 0xffff)) < (unsigned)p.W;
            if (k * PL + PL <= kHaloPix || pl + k * PL < kHaloPix)
              cp_async16(dst + (uint32_t)(k * PL) * 16u, ok ? orgc + (from_up ? rel_up[k] : rel[k]) : p.in, ok ? 16u : 0u);
          }
          cp_async_arrive_noinc(afull_bar(astage));
          if (++astage == p.a_stages) { astage = 0; aphase ^= 1; }
        }
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
    }
  }
  if (!HALO && warp == 0) {
    // ================================ TMA producer (whole warp converged, one elected lane issues) ================================
    int stage = 0; uint32_t phase = 0;
    const int pad = p.ksize >> 1;
    for (int w = blockIdx.x; w < p.num_work; w += gridDim.x) {
      const WorkItem it = decode_work(p, w);
      int brow = it.group * taps * p.cout_slab + it.n_tile * p.block_n + it.img * p.b_img_rows;
      int dy = -pad, dx = -pad, cb = 0;
      for (int k = 0; k < kiters; ++k) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        if (SPLIT) {
          // virtual K block cb = part * nblk_phys + physical block: A_hi.W_hi, A_lo.W_hi, A_hi.W_lo
          if (elect_one()) {
            const uint32_t sa = stage_base + stage * p.stage_bytes, sb = sa + Cfg::kABytes;
            const int s3 = cb / p.nblk_phys, cbp = cb - s3 * p.nblk_phys;
            mbar_expect_tx(full_bar(stage), Cfg::kABytes + b_bytes);
            tma_load_5d(sa, &tmA, full_bar(stage), cbp * BK, s3 == 1 ? 1 : 0, it.x0 + dx, it.y0 + dy, it.img);
            tma_load_2d(sb, &tmB, full_bar(stage), cbp * BK + (s3 == 2 ? p.cin_pad1 : 0), brow);
          }
        } else if (elect_one()) {
          const uint32_t sa = stage_base + stage * p.stage_bytes, sb = sa + Cfg::kABytes;
          if (pair) {
            // both CTAs' boxes complete on the leader's barrier; the leader expects the bytes of both
            const uint32_t lead_full = mapa_shared(full_bar(stage), 0);
            if (cta_rank == 0) mbar_expect_tx(full_bar(stage), 2u * (Cfg::kABytes + b_bytes));
            tma_load_4d_2sm(sa, &tmA, lead_full, cb * BK, it.x0 + dx, it.y0 + dy, it.img);
            tma_load_2d_2sm(sb, &tmB, lead_full, cb * BK, brow + (int)cta_rank * (p.block_n >> 1));
          } else {
            mbar_expect_tx(full_bar(stage), Cfg::kABytes + b_bytes);
            tma_load_4d(sa, &tmA, full_bar(stage), cb * BK, it.x0 + dx, it.y0 + dy, it.img);
            tma_load_2d(sb, &tmB, full_bar(stage), cb * BK, brow);
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
        if (++cb == p.kblocks_per_tap) { cb = 0; brow += p.cout_slab; if (++dx > pad) { dx = -pad; ++dy; } }
      }
    }
  } else if (!HALO && warp == 1) {
    // ================================ MMA issuer (whole warp converged, one elected lane issues) ================================
    const uint32_t idesc = make_idesc_f16(p.block_n, pair ? 2 * kBlockM : kBlockM);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    // descriptors = a constant high word | the 14-bit (address >> 4) that advances by adds (no per-iteration multiply / shift / mask)
    const uint64_t desc_hi = make_kmajor_desc<BK>(0);
    const uint32_t a_units0 = stage_base >> 4, stage_units = (uint32_t)p.stage_bytes >> 4;
    constexpr uint32_t kBOffUnits = Cfg::kABytes >> 4;
    uint32_t a_units = a_units0;
    // pair: the leader issues for both CTAs (its work items and the peer's advance in lock step); the peer's MMA warp idles
    for (int w = blockIdx.x; w < p.num_work && cta_rank == 0; w += gridDim.x) {
      mbar_wait(tempty_bar(acc), acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_stride);
      for (int k = 0; k < kiters; ++k) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = desc_hi | a_units, bdesc = desc_hi | (a_units + kBOffUnits);
          umma_x<pair>(d_tmem, adesc, bdesc, idesc, k ? 1u : 0u);
#pragma unroll
          for (int kk = 1; kk < BK / 16; ++kk) umma_acc_x<pair>(d_tmem, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc);
          umma_commit_x<pair>(empty_bar(stage));      // frees the smem slot (of both CTAs of a pair) when these MMAs retire
          if (k == kiters - 1) umma_commit_x<pair>(tfull_bar(acc));
        }
        __syncwarp();
        a_units += stage_units;
        if (++stage == kStages) { stage = 0; phase ^= 1; a_units = a_units0; }
      }
      if (++acc == n_acc) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ================================ epilogue (two groups of 4 warps, group g drains accumulator g) ================================
    const int g = (warp - 4) >> 2;
    const int te = (threadIdx.x - 128) & 127;   // 0..127 == output row of the tile == TMEM lane
    const int q = warp & 3;                     // TMEM lane quarter this warp may touch
    const bool issuer_warp = q == 0;            // first warp of the group issues the group's TMA traffic
    const int nchunks = (p.block_n + kChunkC - 1) / kChunkC;
    const uint32_t stg0 = staging_base + g * 2 * kStagingBytes;
    uint32_t cc = 0, ac = 0;                      // staging / export tile parity counters
    uint32_t res_phase = 0;                       // bit b: parity of the next residual load into staging buffer b
    // the CTA's i-th tile uses accumulator i % n_acc (phase (i / n_acc) & 1) and is drained by group i % epi_groups
    const int G = p.epi_groups;
    int ti = g, acc = g % n_acc; uint32_t acc_phase = (uint32_t)(g / n_acc) & 1u;
    const bool ts_on = (p.debug & 16) && p.dbg_ts && blockIdx.x == 0 && threadIdx.x == 128;
    int ts_tile = 0;
#define HIS_TS(k) do { if (ts_on && ts_tile < 64) p.dbg_ts[ts_tile * 32 + (k)] = (unsigned long long)clock64(); } while (0)
    for (int w = blockIdx.x + g * gridDim.x; w < p.num_work; w += G * gridDim.x) {
      const WorkItem it = decode_work(p, w);
      HIS_TS(0);
      const uint32_t acc_col = (uint32_t)(acc * p.acc_stride);
      const int chbase = it.n_tile * p.block_n;
      const float* shp = p.n_tiles == 1 ? s_shift : p.shift + chbase;
      // A chunk goes through the swizzled staging buffer + TMA when its 32-channel x 128-pixel box lies fully inside the tensor.
      // Clipped boxes (channel tail, image edge) take a far slower path inside the TMA unit (measured: a half-clipped store box
      // costs ~20 full ones), so those chunks are read / written straight from registers, 16 bytes per 8 channels.
      const bool tile_full = it.y0 + p.bh <= p.H && it.x0 + p.bw <= p.W;
      const int ntma = !p.direct_ok ? nchunks : (!tile_full && p.direct_ok > 1) ? 0 : min(nchunks, max(0, (p.cout - chbase) / kChunkC));
      // direct_ok == 3: the residual of an edge-clipped tile is read straight from global memory (a clipped residual TMA box
      // is served far slower than a full one), the output still leaves through the staging buffer + TMA store
      const bool res_glob = RES && !tile_full && p.direct_ok == 3;
      if (RES && !res_glob && issuer_warp && elect_one()) {  // prefetch the first two residual chunks while the MMAs run
        tma_wait_read<0>();
        if (SPLIT) {       // a chunk owns both staging buffers of the group (hi plane, lo plane): prefetch the first chunk's pair
          if (ntma > 0) {
            mbar_expect_tx(res_bar(2 * g), 2 * kStagingBytes);
            tma_load_5d(stg0, &tmR, res_bar(2 * g), chbase, 0, it.x0, it.y0, it.img);
            tma_load_5d(stg0 + kStagingBytes, &tmR, res_bar(2 * g), chbase, 1, it.x0, it.y0, it.img);
          }
        } else {
          for (int j = 0; j < 2 && j < ntma; ++j) {
            const int b = (cc + j) & 1;
            mbar_expect_tx(res_bar(2 * g + b), kStagingBytes);
            tma_load_4d(stg0 + b * kStagingBytes, &tmR, res_bar(2 * g + b), chbase + j * kChunkC, it.x0, it.y0, it.img);
          }
        }
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      HIS_TS(1);
      float tacc0 = 0.0f, tacc1 = 0.0f;
      constexpr bool TAIL = EPI == EPI_TAIL;
      const bool store_main = (!TAIL || p.store_main) && !(p.debug & 1);
      const int py = it.y0 + te / p.bw, px = it.x0 + te % p.bw;
      const bool inb = py < p.H && px < p.W;
      const long long pix = ((long long)it.img * p.H + py) * p.W + px;
      float* aux_px = nullptr;
      if (EPI == EPI_AUX && inb) aux_px = p.aux_out + ((long long)it.img * p.cout * p.H + py) * p.W + px;
      const float rs = (p.row_scale && inb) ? __ldg(p.row_scale + pix) : 1.0f;
      // per-(image, channel) residual scale: the image's row is staged in shared memory once per tile (it was 8 global loads per
      // 8 channels per thread before: the epilogue of the mask-resolution layer went from 2.6 to 4.3 ms)
      float* s_rsc = s_shift + 256 + g * 256;
      const bool has_rsc = RES && p.res_scale != nullptr;
      if (has_rsc) {
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
        const float* rsrc = p.res_scale + (long long)it.img * p.cout + chbase;
        for (int i = te; i < p.block_n; i += 128) s_rsc[i] = (chbase + i < p.cout) ? __ldg(rsrc + i) : 0.0f;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
      }
      // several N tiles: the tile's slice of the shift vector goes to the group's (free) residual-scale row
      const bool shift_stage = p.n_tiles > 1 && !has_rsc && !p.ln_partials && !(p.debug & 32);
      if (shift_stage) {
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
        for (int i = te; i < p.block_n; i += 128) s_rsc[i] = __ldg(p.shift + chbase + i);
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
      }
      const uint32_t sha0 = shift_stage ? shift_base + (uint32_t)(1 + g) * 1024u : shift_base;
      float st_sum = 0.0f, st_max = -INFINITY;
      float ln_s = 0.0f, ln_q = 0.0f;
      // Straight-line chunk body for the common case (plain fp16 store of a full 32-channel chunk through the staging buffer, no
      // per-pixel statistics / LayerNorm sums / residual scale): one warp per scheduler runs this code, so what bounds a chunk is
      // the dependent-issue latency of its instruction stream -- the general body below is four branchy 8-channel blocks the
      // compiler cannot interleave (measured 1 150 - 2 000 clk per chunk of a 256-wide layer against ~110 clk for the TMEM load).
      const bool fast_tile = (p.n_tiles == 1 || shift_stage) && !(p.ln_partials && ((p.phase_merge > 1 && p.phase_slab != p.cout) || (!p.direct_ok && (p.cout % kChunkC) != 0))) && !res_glob && !(p.debug & 32) &&
                             !(has_rsc && (p.debug & 128)) &&
                             (EPI != EPI_TAIL || tail_smem) && (EPI != EPI_AUX || p.aux_tma);
      // residual operand (not SPLIT): the load of chunk j+1 is issued from the middle of chunk j (see below) instead of the top of j+1
      // (needs a group barrier inside every staged chunk so that no warp runs ahead of the group by more than a chunk: the one
      // before the chunk's store, or -- a tail-only layer, store_main == 0, has no store -- the one kept at the top of its chunks)
      const bool res_early = RES && !SPLIT && !res_glob && !(p.debug & 64);
      HIS_TS(2);
      for (int j = 0; j < nchunks; ++j) {
        const bool direct = j >= ntma;
        const int tsb = 3 + (j < 3 ? j : 2) * 7;
        HIS_TS(tsb);
        const int b = SPLIT ? 0 : (cc & 1);     // SPLIT: staging[0] = hi plane, staging[1] = lo plane of the chunk
        const uint32_t stg = stg0 + b * kStagingBytes;
        const int cl0 = j * kChunkC;            // first channel of this chunk within the N tile
        const int ch0 = chbase + cl0;           // ... and within the layer
        const int ncol = min(kChunkC, p.block_n - cl0);   // valid accumulator columns in this chunk (16 or 32)
        const bool aux_tma = EPI == EPI_AUX && p.aux_tma;
        // the export tile this chunk writes (and staging[b]) are free again: with two tiles the previous chunk's stores may still drain
        if (aux_tma && issuer_warp && elect_one()) { if (p.aux_bufs > 1) tma_wait_read<1>(); else tma_wait_read<0>(); }
        const uint32_t aux_tile = aux_base + (uint32_t)(g * p.aux_bufs + (p.aux_bufs > 1 ? (int)(ac & 1u) : 0)) * kAuxStagingBytes;
        if (direct && aux_tma) asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
        // Staging hand-over.  lean: every chunk's issuer drains the earlier stores BEFORE the barrier that precedes its own store
        // (see below; residual chunks: before the next chunk's residual load), so after that barrier the other staging buffer is
        // known free to the whole group and a chunk starts without a wait and without a barrier.  The split / export kernels
        // keep the wait + barrier at the top of the chunk.
        const bool lean_sync = !SPLIT && !aux_tma && store_main && !(p.debug & 64);
        if (!direct && !lean_sync) {
          if (SPLIT) {
            if ((!RES || res_glob || j >= 1) && issuer_warp && elect_one()) {
              tma_wait_read<0>();                  // both stores of the previous chunk have drained
              if (RES && !res_glob) {
                mbar_expect_tx(res_bar(2 * g), 2 * kStagingBytes);
                tma_load_5d(stg0, &tmR, res_bar(2 * g), ch0, 0, it.x0, it.y0, it.img);
                tma_load_5d(stg0 + kStagingBytes, &tmR, res_bar(2 * g), ch0, 1, it.x0, it.y0, it.img);
              }
            }
          } else if ((!RES || res_glob || (j >= 2 && !res_early)) && issuer_warp && elect_one()) {
            tma_wait_read<1>();                    // the store that last read staging[b] has drained
            if (RES && !res_glob) {
              mbar_expect_tx(res_bar(2 * g + b), kStagingBytes);
              tma_load_4d(stg, &tmR, res_bar(2 * g + b), ch0, it.x0, it.y0, it.img);
            }
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
        }
        HIS_TS(tsb + 1);
        uint32_t v[kChunkC];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc_col + (uint32_t)cl0;
        if (ncol > 16) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
        tmem_ld_wait();
        HIS_TS(tsb + 2);
        if (j == nchunks - 1) {                  // accumulator fully read -> hand TMEM back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (pair) mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));   // the leader's MMA thread waits for both CTAs
            else mbar_arrive(tempty_bar(acc));
          }
        }
        if (RES && !direct && !res_glob) { mbar_wait(res_bar(2 * g + b), (res_phase >> b) & 1u); res_phase ^= 1u << b; }
        if (res_early && j >= 1 && j + 1 < ntma && issuer_warp && elect_one()) {
          // chunk j+1's residual -> the other staging buffer, whose last reader is the store of chunk j-1 (issued a whole chunk
          // ago: the wait is short); the load then has this chunk's math and store to land instead of being waited for at once
          tma_wait_read<0>();
          mbar_expect_tx(res_bar(2 * g + (b ^ 1)), kStagingBytes);
          tma_load_4d(stg0 + (b ^ 1) * kStagingBytes, &tmR, res_bar(2 * g + (b ^ 1)), ch0 + kChunkC, it.x0, it.y0, it.img);
        }
        HIS_TS(tsb + 3);
        uint8_t* row_ptr = smem_gen + (stg - smem_base) + te * (kChunkC * 2);
        float* aux_row = reinterpret_cast<float*>(smem_gen + (aux_tile - smem_base)) + te;
        const bool fast = fast_tile && !direct && ncol == kChunkC;
        if (fast) {
          const uint32_t rowa = stg + (uint32_t)te * (kChunkC * 2), sw = (uint32_t)(te >> 1) & 3u;     // SWIZZLE_64B row of this pixel
          const uint32_t sha = sha0 + (uint32_t)cl0 * 4u;
          const uint32_t rsca = shift_base + (uint32_t)(1 + g) * 1024u + (uint32_t)cl0 * 4u;
          // two halves of 16 channels: all shared-memory loads of a half first (the volatile accesses keep their program order, so
          // loads placed after a store would wait for it), then the math, then the stores
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint4 rv[2], rl[2], sv[4];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              if (RES) rv[k] = lds128(rowa + (((uint32_t)(2 * h + k) ^ sw) << 4));
              if (RES && SPLIT) rl[k] = lds128(rowa + kStagingBytes + (((uint32_t)(2 * h + k) ^ sw) << 4));      // lo plane of the residual
              sv[2 * k] = lds128(sha + (uint32_t)(2 * h + k) * 32u);
              sv[2 * k + 1] = lds128(sha + (uint32_t)(2 * h + k) * 32u + 16u);
            }
            uint4 ov[2], ol[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int i = 2 * h + k;
              const float sh[8] = {__uint_as_float(sv[2 * k].x), __uint_as_float(sv[2 * k].y), __uint_as_float(sv[2 * k].z), __uint_as_float(sv[2 * k].w),
                                   __uint_as_float(sv[2 * k + 1].x), __uint_as_float(sv[2 * k + 1].y), __uint_as_float(sv[2 * k + 1].z),
                                   __uint_as_float(sv[2 * k + 1].w)};
              const __half2* rh = reinterpret_cast<const __half2*>(&rv[k]);
              const __half2* rlh = reinterpret_cast<const __half2*>(&rl[k]);
              __half2* o = reinterpret_cast<__half2*>(&ov[k]);
              __half2* olo = reinterpret_cast<__half2*>(&ol[k]);
              uint4 qa = make_uint4(0u, 0u, 0u, 0u), qb = qa;          // residual scale of the 8 channels (ChannelAttention gate)
              if (RES && has_rsc) { qa = lds128(rsca + (uint32_t)i * 32u); qb = lds128(rsca + (uint32_t)i * 32u + 16u); }
              float y[8];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float t0 = fmaf(__uint_as_float(v[i * 8 + 2 * e]), rs, sh[2 * e]);
                float t1 = fmaf(__uint_as_float(v[i * 8 + 2 * e + 1]), rs, sh[2 * e + 1]);
                float2 r = make_float2(0.0f, 0.0f);
                if (RES) r = __half22float2(rh[e]);
                if (RES && SPLIT) { const float2 q2 = __half22float2(rlh[e]); r.x += q2.x; r.y += q2.y; }
                if (RES && has_rsc) { r.x *= __uint_as_float(e < 2 ? (e ? qa.z : qa.x) : (e == 2 ? qb.x : qb.z)); r.y *= __uint_as_float(e < 2 ? (e ? qa.w : qa.y) : (e == 2 ? qb.y : qb.w)); }
                if (RES == HIS_RES_ADD) { t0 += r.x; t1 += r.y; }
                t0 = epi_act<ACTC>(t0, p); t1 = epi_act<ACTC>(t1, p);
                if (EPI == EPI_AUX) {      // [channel][pixel] fp32 tile: lanes write consecutive words
                  sts32f(aux_tile + (uint32_t)((i * 8 + 2 * e) * kBlockM + te) * 4u, t0);
                  sts32f(aux_tile + (uint32_t)((i * 8 + 2 * e + 1) * kBlockM + te) * 4u, t1);
                }
                if (RES == HIS_RES_MUL) { t0 *= r.x; t1 *= r.y; }
                y[2 * e] = t0; y[2 * e + 1] = t1;
                o[e] = __floats2half2_rn(t0, t1);
                if (SPLIT) { const float2 hh = __half22float2(o[e]); olo[e] = __floats2half2_rn(t0 - hh.x, t1 - hh.y); }     // lo = fp16(y - hi)
              }
              if (p.stats_out) {        // per-pixel channel mean / max (every channel of a full chunk is a real one)
#pragma unroll
                for (int e = 0; e < 8; ++e) { st_sum += y[e]; st_max = fmaxf(st_max, y[e]); }
              }
              if (p.ln_partials && inb) {      // LayerNorm2d sums over the fp16-rounded values the normalise pass will read
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 hh = SPLIT ? make_float2(y[2 * e], y[2 * e + 1]) : __half22float2(o[e]);      // split: hi + lo ~ y itself
                  ln_s += hh.x; ln_q = fmaf(hh.x, hh.x, ln_q);
                  ln_s += hh.y; ln_q = fmaf(hh.y, hh.y, ln_q);
                }
              }
              if (TAIL) {
                const uint4 wa = lds128(sha + kTailRow + (uint32_t)i * 32u), wb = lds128(sha + kTailRow + (uint32_t)i * 32u + 16u);
                const float4 w0 = make_float4(__uint_as_float(wa.x), __uint_as_float(wa.y), __uint_as_float(wa.z), __uint_as_float(wa.w));
                const float4 w1 = make_float4(__uint_as_float(wb.x), __uint_as_float(wb.y), __uint_as_float(wb.z), __uint_as_float(wb.w));
                tacc0 += y[0] * w0.x + y[1] * w0.y + y[2] * w0.z + y[3] * w0.w + y[4] * w1.x + y[5] * w1.y + y[6] * w1.z + y[7] * w1.w;
                if (p.tail_c > 1) {
                  const uint4 ua = lds128(sha + kTailRow + 1024u + (uint32_t)i * 32u), ub = lds128(sha + kTailRow + 1024u + (uint32_t)i * 32u + 16u);
                  const float4 u0 = make_float4(__uint_as_float(ua.x), __uint_as_float(ua.y), __uint_as_float(ua.z), __uint_as_float(ua.w));
                  const float4 u1 = make_float4(__uint_as_float(ub.x), __uint_as_float(ub.y), __uint_as_float(ub.z), __uint_as_float(ub.w));
                  tacc1 += y[0] * u0.x + y[1] * u0.y + y[2] * u0.z + y[3] * u0.w + y[4] * u1.x + y[5] * u1.y + y[6] * u1.z + y[7] * u1.w;
                }
              }
            }
            if (store_main) {
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                sts128(rowa + (((uint32_t)(2 * h + k) ^ sw) << 4), ov[k]);
                if (SPLIT) sts128(rowa + kStagingBytes + (((uint32_t)(2 * h + k) ^ sw) << 4), ol[k]);
              }
            }
          }
        }
#pragma unroll
        for (int i = 0; i < kChunkC / 8; ++i) {
          if (!fast && i * 8 < ncol) {
            const int cl = cl0 + i * 8, c = ch0 + i * 8;
            uint4* cell = reinterpret_cast<uint4*>(row_ptr + ((i ^ ((te >> 1) & 3)) << 4));   // SWIZZLE_64B
            uint4* cell_lo = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(cell) + kStagingBytes);      // SPLIT: lo plane
            float r[8];
            if (RES) {
              uint4 rv = make_uint4(0u, 0u, 0u, 0u);
              if (!direct && !res_glob) rv = *cell;
              else if (inb && c + 8 <= p.cout) rv = __ldg(reinterpret_cast<const uint4*>(p.res + pix * p.res_cs + c));
              else if (inb) {
                __half* rh1 = reinterpret_cast<__half*>(&rv);
                for (int e = 0; e < 8; ++e) if (c + e < p.cout) rh1[e] = p.res[pix * p.res_cs + c + e];
              }
              const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
              for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(rh[e]); r[2 * e] = f.x; r[2 * e + 1] = f.y; }
              if (SPLIT) {
                uint4 rl = make_uint4(0u, 0u, 0u, 0u);
                if (!direct && !res_glob) rl = *cell_lo;
                else if (inb && c + 8 <= p.cout) rl = __ldg(reinterpret_cast<const uint4*>(p.res + pix * p.res_cs + p.res_lo + c));
                else if (inb) {
                  __half* rh1 = reinterpret_cast<__half*>(&rl);
                  for (int e = 0; e < 8; ++e) if (c + e < p.cout) rh1[e] = p.res[pix * p.res_cs + p.res_lo + c + e];
                }
                const __half2* rl2 = reinterpret_cast<const __half2*>(&rl);
#pragma unroll
                for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(rl2[e]); r[2 * e] += f.x; r[2 * e + 1] += f.y; }
              }
              if (has_rsc) {
                const float4 q0 = *reinterpret_cast<const float4*>(s_rsc + cl), q1 = *reinterpret_cast<const float4*>(s_rsc + cl + 4);
                r[0] *= q0.x; r[1] *= q0.y; r[2] *= q0.z; r[3] *= q0.w; r[4] *= q1.x; r[5] *= q1.y; r[6] *= q1.z; r[7] *= q1.w;
              }
            }
            const float4 t0 = *reinterpret_cast<const float4*>(shp + cl), t1 = *reinterpret_cast<const float4*>(shp + cl + 4);
            const float sh[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
            float y[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float t = __uint_as_float(v[i * 8 + e]) * rs + sh[e];
              if (RES == HIS_RES_ADD) t += r[e];
              t = epi_act<ACTC>(t, p);
              if (EPI == EPI_AUX) {
                if (aux_tma) aux_row[(i * 8 + e) * kBlockM] = t;      // [channel][pixel]: lanes write consecutive words
                else if (aux_px && c + e < p.cout) aux_px[(long long)(c + e) * p.H * p.W] = t;
              }
              if (RES == HIS_RES_MUL) t *= r[e];
              y[e] = t;
            }
            if (p.stats_out) {
#pragma unroll
              for (int e = 0; e < 8; ++e) if (c + e < p.cout) { st_sum += y[e]; st_max = fmaxf(st_max, y[e]); }
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
              if (p.ln_partials && inb) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  // SPLIT stores hi + lo ~ y itself; plain fp16 stores the rounded value
                  const float2 h = SPLIT ? make_float2(y[2 * e], y[2 * e + 1]) : __half22float2(o[e]);
                  const int cv = (p.phase_merge > 1 ? c % p.phase_slab : c) + 2 * e;      // channel within its ConvT phase
                  if (cv < p.cout) { ln_s += h.x; ln_q = fmaf(h.x, h.x, ln_q); }
                  if (cv + 1 < p.cout) { ln_s += h.y; ln_q = fmaf(h.y, h.y, ln_q); }
                }
              }
              if (!direct) *cell = *reinterpret_cast<uint4*>(o);
              else if (inb && c + 8 <= p.cout) *reinterpret_cast<uint4*>(p.out + pix * p.out_cs + c) = *reinterpret_cast<uint4*>(o);
              else if (inb) {
                const __half* oh = reinterpret_cast<const __half*>(o);
                for (int e = 0; e < 8; ++e) if (c + e < p.cout) p.out[pix * p.out_cs + c + e] = oh[e];
              }
              if (SPLIT) {       // lo = fp16(y - hi): the pair carries ~21 significant bits
                __half2 ol[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 h = __half22float2(o[e]);
                  ol[e] = __floats2half2_rn(y[2 * e] - h.x, y[2 * e + 1] - h.y);
                }
                if (!direct) *cell_lo = *reinterpret_cast<uint4*>(ol);
                else if (inb && c + 8 <= p.cout) *reinterpret_cast<uint4*>(p.out + pix * p.out_cs + p.out_lo + c) = *reinterpret_cast<uint4*>(ol);
                else if (inb) {
                  const __half* oh = reinterpret_cast<const __half*>(ol);
                  for (int e = 0; e < 8; ++e) if (c + e < p.cout) p.out[pix * p.out_cs + p.out_lo + c + e] = oh[e];
                }
              }
            }
          }
        }
        HIS_TS(tsb + 4);
        if (lean_sync && store_main && !direct && !(RES && !res_glob) && issuer_warp && elect_one()) tma_wait_read<0>();
        if ((store_main && !direct) || aux_tma) {
          fence_proxy_async();
          asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
          HIS_TS(tsb + 5);
          if (issuer_warp && elect_one()) {
            if (store_main && !direct) {
              // merged ConvT phases: the chunk belongs to phase (group * merge + cl0 / phase_slab) and to its output map
              const int ph = p.phase_merge > 1 ? cl0 / p.phase_slab : 0;
              const int ch_st = p.phase_merge > 1 ? cl0 - ph * p.phase_slab : ch0;
              const int gsel = p.phase_merge > 1 ? it.group * p.phase_merge + ph : it.group;
              const CUtensorMap* tmOj = gsel == 0 ? &tmO0 : gsel == 1 ? &tmO1 : gsel == 2 ? &tmO2 : &tmO3;
              if (SPLIT) {
                tma_store_5d(tmOj, stg0, ch_st, 0, it.x0, it.y0, it.img);
                tma_store_5d(tmOj, stg0 + kStagingBytes, ch_st, 1, it.x0, it.y0, it.img);
              } else {
                tma_store_4d(tmOj, stg, ch_st, it.x0, it.y0, it.img);
              }
            }
            // fp32 NCHW export: box {bw, bh, 32 channels, 1} of the map {W, H, C, N}; the TMA unit clips image edges / channel tail
            if (aux_tma) tma_store_4d(&tmAux, aux_tile, it.x0, it.y0, ch0, it.img);
            tma_commit();
          }
        }
        HIS_TS(tsb + 6);
        if (!direct) ++cc;
        if (aux_tma) ++ac;
      }
      HIS_TS(24);
      ++ts_tile;
      if (p.stats_out && inb) {
        p.stats_out[pix * 2] = st_sum / (float)p.cout;
        p.stats_out[pix * 2 + 1] = st_max;
      }
      if (p.ln_partials) {      // fixed-order reduction: lanes (shuffle tree), then the group's four warps through shared memory
#pragma unroll
        for (int o = 16; o; o >>= 1) { ln_s += __shfl_xor_sync(0xffffffffu, ln_s, o); ln_q += __shfl_xor_sync(0xffffffffu, ln_q, o); }
        float* red = s_shift + 256 + g * 256;          // the residual-scale row of the group (LayerNorm layers carry none)
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
        if (lane == 0) { red[2 * q] = ln_s; red[2 * q + 1] = ln_q; }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
        if (te == 0) {
          p.ln_partials[2 * (long long)w] = ((double)red[0] + (double)red[2]) + ((double)red[4] + (double)red[6]);
          p.ln_partials[2 * (long long)w + 1] = ((double)red[1] + (double)red[3]) + ((double)red[5] + (double)red[7]);
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
      ti += G;
      acc = ti % n_acc; acc_phase = (uint32_t)(ti / n_acc) & 1u;
    }
    if (issuer_warp && elect_one()) tma_wait_all();
#undef HIS_TS
  }

  tc_fence_before();
  __syncthreads();
  if (pair) cluster_sync_all();        // no CTA leaves (or frees TMEM) while its peer can still signal it
  if (warp == 2) {
    tc_fence_after();
    if (pair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

typedef void (*ConvGemmKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                               const CUtensorMap, const CUtensorMap, const CUtensorMap, const ConvGemmParams);
}  // namespace
