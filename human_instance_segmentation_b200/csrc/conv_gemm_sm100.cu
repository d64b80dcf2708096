#include "conv_gemm_sm100_kernel.cuh"

namespace {

template <int BK, int ACTC, bool HALO>
ConvGemmKernel pick_res(int res) {
  return res == 0 ? conv_gemm_sm100_kernel<BK, ACTC, 0, EPI_PLAIN, HALO> : res == 1 ? conv_gemm_sm100_kernel<BK, ACTC, 1, EPI_PLAIN, HALO>
                                                                                  : conv_gemm_sm100_kernel<BK, ACTC, 2, EPI_PLAIN, HALO>;
}
template <int BK, bool HALO>
ConvGemmKernel pick_act(int actc, int res) {
  return actc == 0 ? pick_res<BK, 0, HALO>(res) : actc == 1 ? pick_res<BK, 1, HALO>(res) : pick_res<BK, 2, HALO>(res);
}
template <bool HALO>
ConvGemmKernel pick_bk_kernel(int bk, int actc, int res) {
  return bk == 64 ? pick_act<64, HALO>(actc, res) : bk == 32 ? pick_act<32, HALO>(actc, res) : pick_act<16, HALO>(actc, res);
}
ConvGemmKernel pick_kernel(int bk, int actc, int res, bool halo) {
  return halo ? pick_bk_kernel<true>(bk, actc, res) : pick_bk_kernel<false>(bk, actc, res);
}
// fused-tail variants exist for the shapes that need them: clamp activations (none/relu), no residual or residual-add
template <bool HALO>
ConvGemmKernel pick_tail_h(int bk, int res) {
  if (bk == 64) return res == 0 ? conv_gemm_sm100_kernel<64, ACTC_CLAMP, 0, EPI_TAIL, HALO> : conv_gemm_sm100_kernel<64, ACTC_CLAMP, 1, EPI_TAIL, HALO>;
  if (bk == 32) return res == 0 ? conv_gemm_sm100_kernel<32, ACTC_CLAMP, 0, EPI_TAIL, HALO> : conv_gemm_sm100_kernel<32, ACTC_CLAMP, 1, EPI_TAIL, HALO>;
  return res == 0 ? conv_gemm_sm100_kernel<16, ACTC_CLAMP, 0, EPI_TAIL, HALO> : conv_gemm_sm100_kernel<16, ACTC_CLAMP, 1, EPI_TAIL, HALO>;
}
ConvGemmKernel pick_tail_kernel(int bk, int res, bool halo) { return halo ? pick_tail_h<true>(bk, res) : pick_tail_h<false>(bk, res); }
// fp32 NCHW export variants: BK = 64 layers only (the 256-channel trunk / gate layers whose outputs the reference returns)
template <int ACTC, bool HALO>
ConvGemmKernel pick_aux_res(int res) {
  return res == 0 ? conv_gemm_sm100_kernel<64, ACTC, 0, EPI_AUX, HALO> : res == 1 ? conv_gemm_sm100_kernel<64, ACTC, 1, EPI_AUX, HALO>
                                                                                 : conv_gemm_sm100_kernel<64, ACTC, 2, EPI_AUX, HALO>;
}
template <bool HALO>
ConvGemmKernel pick_aux_h(int actc, int res) {
  return actc == 0 ? pick_aux_res<0, HALO>(res) : actc == 1 ? pick_aux_res<1, HALO>(res) : pick_aux_res<2, HALO>(res);
}
ConvGemmKernel pick_aux_kernel(int actc, int res, bool halo) { return halo ? pick_aux_h<true>(actc, res) : pick_aux_h<false>(actc, res); }
// CTA-pair variants (BK = 64; per-tap TMA mode or halo mode): launched as clusters of two
template <int ACTC, bool HALO>
ConvGemmKernel pick_pair_res(int res) {
  return res == 0 ? conv_gemm_sm100_kernel<64, ACTC, 0, EPI_PLAIN, HALO, true> : res == 1 ? conv_gemm_sm100_kernel<64, ACTC, 1, EPI_PLAIN, HALO, true>
                                                                                      : conv_gemm_sm100_kernel<64, ACTC, 2, EPI_PLAIN, HALO, true>;
}
template <bool HALO>
ConvGemmKernel pick_pair_h(int actc, int res) {
  return actc == 0 ? pick_pair_res<0, HALO>(res) : actc == 1 ? pick_pair_res<1, HALO>(res) : pick_pair_res<2, HALO>(res);
}
ConvGemmKernel pick_pair_kernel(int actc, int res, bool halo) { return halo ? pick_pair_h<true>(actc, res) : pick_pair_h<false>(actc, res); }
ConvGemmKernel pick_pair_tail_kernel(int res, bool halo) {
  if (halo) return res == 0 ? conv_gemm_sm100_kernel<64, ACTC_CLAMP, 0, EPI_TAIL, true, true> : conv_gemm_sm100_kernel<64, ACTC_CLAMP, 1, EPI_TAIL, true, true>;
  return res == 0 ? conv_gemm_sm100_kernel<64, ACTC_CLAMP, 0, EPI_TAIL, false, true> : conv_gemm_sm100_kernel<64, ACTC_CLAMP, 1, EPI_TAIL, false, true>;
}
template <int ACTC, bool HALO>
ConvGemmKernel pick_pair_aux_res(int res) {
  return res == 0 ? conv_gemm_sm100_kernel<64, ACTC, 0, EPI_AUX, HALO, true> : res == 1 ? conv_gemm_sm100_kernel<64, ACTC, 1, EPI_AUX, HALO, true>
                                                                                    : conv_gemm_sm100_kernel<64, ACTC, 2, EPI_AUX, HALO, true>;
}
template <bool HALO>
ConvGemmKernel pick_pair_aux_h(int actc, int res) {
  return actc == 0 ? pick_pair_aux_res<0, HALO>(res) : actc == 1 ? pick_pair_aux_res<1, HALO>(res) : pick_pair_aux_res<2, HALO>(res);
}
ConvGemmKernel pick_pair_aux_kernel(int actc, int res, bool halo) { return halo ? pick_pair_aux_h<true>(actc, res) : pick_pair_aux_h<false>(actc, res); }
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
                   int box_c, int box_w, int box_h, int lo = 0) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return his_set_error(HIS_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  if (lo > 0) {
    // split-fp16 activation: 5-D map {C, part (hi / lo), W, H, N}; the lo plane of a pixel lies `lo` elements after its hi plane
    cuuint64_t dims[5] = {(cuuint64_t)C, 2, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)lo * 2, (cuuint64_t)sW * 2, (cuuint64_t)sH * 2, (cuuint64_t)sN * 2};
    cuuint32_t box[5] = {(cuuint32_t)box_c, 1, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    if (((uintptr_t)base & 15) || (strides[0] & 15) || (strides[1] & 15) || (strides[2] & 15) || (strides[3] & 15))
      return his_set_error(HIS_ERR_INVALID_ARG, "split activation slice is not 16-byte aligned (offsets / strides must be multiples of 8)");
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(box_c), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      char buf[128]; snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(split act) failed: %d", (int)r);
      return his_set_error(HIS_ERR_DRIVER, buf);
    }
    return HIS_OK;
  }
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
  CUtensorMap tmA, tmB, tmO[4], tmR, tmAux;
  ConvGemmParams p;
  int grid, bk, smem, actc, res_mode, transposed, cin_pad, halo, b_box_rows, pair_grid_checked, split;
  int k1_split, num_sms;        // k1_split: a 256-wide 1x1 layer running as two 128-wide N tiles (see his_conv_gemm_create)
  const void* w_packed;
  ConvGemmKernel kernel;
};

// K-block width: the largest of {64,32,16} whose padded K stays within 25 % of the best padding
int pick_bk(int cin) {
  const int c16 = (cin + 15) / 16 * 16, c32 = (cin + 31) / 32 * 32, c64 = (cin + 63) / 64 * 64;
  if (c64 * 4 <= c16 * 5) return 64;
  if (c32 * 4 <= c16 * 5) return 32;
  return 16;
}

// split-fp16 ("strict") instantiations live in conv_gemm_sm100_split.cu (same kernel template, SPLIT = true); the function
// pointers cross the translation units as void* (the parameter structs are textually identical)
}  // namespace
extern "C" void* his_gemm_pick_split_kernel(int bk, int actc, int res, int epi, int halo);
extern "C" int his_gemm_split_set_smem_attr(void);
namespace {

// One-time work PER DEVICE: the > 48 KB dynamic shared memory opt-in is a per-device function attribute, and the SM count sizes
// the persistent grids (a process may drive several GPUs).
constexpr int kMaxDevices = 64;
int g_dev_sms[kMaxDevices] = {0};

int device_init(int* num_sms) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return his_set_error(HIS_ERR_NO_DEVICE, "no CUDA device");
  if (g_dev_sms[dev]) { *num_sms = g_dev_sms[dev]; return HIS_OK; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return his_set_error(HIS_ERR_NO_DEVICE, "no CUDA device");
  if (prop.major != 10) return his_set_error(HIS_ERR_UNSUPPORTED, "conv_gemm_sm100 needs a compute-capability 10.x device (B200)");
  for (int h = 0; h < 2; ++h) {
    for (int bk = 16; bk <= 64; bk *= 2)
      for (int a = 0; a < 3; ++a)
        for (int r = 0; r < 3; ++r)
          if (cudaFuncSetAttribute(pick_kernel(bk, a, r, h), cudaFuncAttributeMaxDynamicSharedMemorySize, smem_for(bk)) != cudaSuccess)
            return his_set_error(HIS_ERR_LAUNCH, "cannot raise dynamic shared memory limit");
    for (int bk = 16; bk <= 64; bk *= 2)
      for (int r = 0; r < 2; ++r)
        if (cudaFuncSetAttribute(pick_tail_kernel(bk, r, h), cudaFuncAttributeMaxDynamicSharedMemorySize, smem_for(bk)) != cudaSuccess)
          return his_set_error(HIS_ERR_LAUNCH, "cannot raise dynamic shared memory limit");
    for (int a = 0; a < 3; ++a)
      for (int r = 0; r < 3; ++r)
        if (cudaFuncSetAttribute(pick_aux_kernel(a, r, h), cudaFuncAttributeMaxDynamicSharedMemorySize, smem_for(64)) != cudaSuccess)
          return his_set_error(HIS_ERR_LAUNCH, "cannot raise dynamic shared memory limit");
  }
  for (int h = 0; h < 2; ++h)
    for (int a = 0; a < 3; ++a)
      for (int r = 0; r < 3; ++r)
        if (cudaFuncSetAttribute(pick_pair_kernel(a, r, h), cudaFuncAttributeMaxDynamicSharedMemorySize, smem_for(64)) != cudaSuccess ||
            cudaFuncSetAttribute(pick_pair_aux_kernel(a, r, h), cudaFuncAttributeMaxDynamicSharedMemorySize, smem_for(64)) != cudaSuccess ||
            (r < 2 && cudaFuncSetAttribute(pick_pair_tail_kernel(r, h), cudaFuncAttributeMaxDynamicSharedMemorySize, smem_for(64)) != cudaSuccess))
          return his_set_error(HIS_ERR_LAUNCH, "cannot raise dynamic shared memory limit");
  if (int rc = his_gemm_split_set_smem_attr()) return rc;
  g_dev_sms[dev] = prop.multiProcessorCount;
  *num_sms = g_dev_sms[dev];
  return HIS_OK;
}

// Widest N tile that runs in halo mode.  Measured on the full B0 step: N <= 96 layers gain 2-3.5x from the halo mode, N = 128
// layers are ahead again in per-tap TMA mode (12.7k vs 12.35k ROI-masks/s), N = 256 is tensor bound in either.
int halo_max_n() {
  if (const char* e = getenv("HIS_GEMM_HALO_MAXN")) return atoi(e);
  return 96;
}

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
                         int ksize, int transposed, int act, float act_beta, int res_mode, int split) {
  if (!out_plan || !in || !w_packed || !out || !shift) return his_set_error(HIS_ERR_INVALID_ARG, "null pointer");
  split = split ? 1 : 0;
  if (split && ((in_cs % 16) || (out_cs % 16) || (res && (res_cs % 16)) || (cin_pad % 16) || cin_pad / 2 < cin))
    return his_set_error(HIS_ERR_INVALID_ARG, "split operands: pixel strides and the packed K ([W_hi | W_lo]) must be multiples of 16");
  if (!(ksize == 1 || ksize == 3) || (transposed && ksize != 1)) return his_set_error(HIS_ERR_UNSUPPORTED, "ksize must be 1 or 3 (transposed: k2s2 packed as 4 1x1 groups)");
  if (res_mode != HIS_RES_NONE && !res) return his_set_error(HIS_ERR_INVALID_ARG, "res_mode set without residual tensor");
  if (res_mode != HIS_RES_NONE && transposed) return his_set_error(HIS_ERR_UNSUPPORTED, "residual with transposed conv");
  if ((in_cs % 8) || (out_cs % 8) || (cin_pad % 8) || cin_pad < cin) return his_set_error(HIS_ERR_INVALID_ARG, "channel strides must be multiples of 8");
  int num_sms = 0;
  if (int rc = device_init(&num_sms)) return rc;
  ConvGemmPlan* pl = new ConvGemmPlan();
  memset(pl, 0, sizeof(*pl));
  ConvGemmParams& p = pl->p;
  p.n_img = n_img; p.H = H; p.W = W; p.ksize = ksize;
  // halo mode: every 3x3 conv whose input slice is 16-byte addressable per 8-channel group (HIS_GEMM_HALO=0 disables it,
  // HIS_GEMM_HALO_MAXN=n keeps the per-tap TMA path for wider N tiles)
  int tn_tiles = 0, tblock_n = 0;
  his_conv_gemm_tile_n(cout, &tn_tiles, &tblock_n);
  int halo = ksize == 3 && !transposed && (cin % 8) == 0;
  if (const char* e = getenv("HIS_GEMM_HALO")) halo = halo && atoi(e) != 0;
  // wide single-N-tile layers run as CTA pairs (cta_group::2) in the per-tap TMA mode; with HIS_GEMM_PAIR_HALO=n those with
  // Cout >= n use the halo window as well (A read once instead of nine times through L2, each CTA streaming half of every tap's
  // weight tile).  Measured equal or slightly slower (128->128 @128x96: 0.985 vs 0.979 ms, 256->256: 1.565 vs 1.519 ms): with the
  // weight traffic halved the pair is no longer L2 bound, so the default keeps the simpler per-tap pair.
  int pair_halo_min_n = 1 << 30;
  if (const char* e = getenv("HIS_GEMM_PAIR_HALO")) pair_halo_min_n = atoi(e) > 0 ? atoi(e) : (1 << 30);
  if (const char* e = getenv("HIS_GEMM_PAIR")) if (atoi(e) <= 0) pair_halo_min_n = 1 << 30;
  const bool pair_halo = !split && halo && tn_tiles == 1 && cin > 32 && tblock_n >= pair_halo_min_n && (tblock_n % 32) == 0 &&
                         ((his_div_up(W, kHaloBw) * his_div_up(H, kHaloBh)) % 2) == 0 && n_img * his_div_up(W, kHaloBw) * his_div_up(H, kHaloBh) >= 2;
  halo = halo && (tblock_n <= halo_max_n() || pair_halo);
  pl->halo = halo;
  // pick the 128-pixel rectangle with the least padded area
  long long best = -1;
  if (halo) { p.bw = kHaloBw; p.bh = kHaloBh; best = 0; }
  for (int bw = 1; bw <= 128 && !halo; bw <<= 1) {
    int bh = 128 / bw;
    long long area = (long long)his_div_up(W, bw) * bw * his_div_up(H, bh) * bh;
    if (best < 0 || area < best || (area == best && bw >= 8 && bw <= 32)) { best = area; p.bw = bw; p.bh = bh; }
  }
  p.tiles_x = his_div_up(W, p.bw); p.tiles_y = his_div_up(H, p.bh);
  const int bk = halo ? (cin > 32 ? 64 : cin > 16 ? 32 : 16) : pick_bk(cin);   // halo mode skips the MMAs of an all-padding K=16 step
  pl->bk = bk;
  // split: three virtual K blocks per physical one (A_hi.W_hi, A_lo.W_hi, A_hi.W_lo); cin_pad counts [W_hi | W_lo]
  p.nblk_phys = his_div_up(cin, bk);
  p.kblocks_per_tap = p.nblk_phys * (split ? 3 : 1);
  p.cin_pad1 = split ? cin_pad / 2 : cin_pad;
  p.in_lo = split ? in_cs / 2 : 0; p.out_lo = split ? out_cs / 2 : 0; p.res_lo = split ? res_cs / 2 : 0; p.up_lo = 0;
  pl->split = split;
  his_conv_gemm_tile_n(cout, &p.n_tiles, &p.block_n);
  // A 1x1 layer with a 256-wide N tile runs as two 128-wide N tiles (same packed weights: the slab stays 256 rows), and a
  // ConvTranspose with 128-wide phases keeps its four phase groups instead of merging two per tile, so that THREE epilogue groups
  // fit the accumulator ring (3 x 128 columns): these layers are bound by the drain rate of two groups (8 chunks of ~1 350 clk per
  // 256-wide tile), not by HBM; A is read twice, the second time from L2.  Measured: ConvT 256->128 0.997 -> 0.811 ms, gate layer
  // 0.644 -> 0.575, 256->256 0.442 -> 0.427.  HIS_GEMM_K1_SPLITN=0 restores the 256-wide tiles.
  int k1_split = 1;
  if (const char* e = getenv("HIS_GEMM_K1_SPLITN")) k1_split = atoi(e);
  pl->k1_split = 0; pl->w_packed = w_packed; pl->num_sms = num_sms;
  if (k1_split && ksize == 1 && !split && p.n_tiles == 1 && p.block_n == 256 && !transposed) { p.n_tiles = 2; p.block_n = 128; pl->k1_split = 1; }
  p.groups = transposed ? 4 : 1;
  p.cout_slab = p.n_tiles * p.block_n;
  p.phase_merge = 1; p.phase_slab = p.block_n;
  if (transposed && p.n_tiles == 1 && (p.block_n % 32) == 0) {
    // ConvTranspose k2s2 = four 1x1 phase GEMMs over the same A: merge as many phases as fit a 256-wide N tile (the packed
    // weight rows of the phases are contiguous, so the slab layout does not change).  HIS_GEMM_CONVT_MERGE=0 keeps four groups.
    int merge = 256 / p.block_n >= 4 ? 4 : 256 / p.block_n >= 2 ? 2 : 1;
    if (const char* e = getenv("HIS_GEMM_CONVT_MERGE")) if (atoi(e) == 0) merge = 1;
    if (k1_split && !split && p.block_n == 128) merge = 1;
    if (merge > 1) {
      p.phase_merge = merge; p.block_n *= merge; p.groups = 4 / merge; p.cout_slab = p.block_n;
    }
  }
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
  p.act_nb2 = p.act_beta * HIS_NEG_LOG2E;
  if (res_mode < 0 || res_mode > 2) { delete pl; return his_set_error(HIS_ERR_INVALID_ARG, "unknown res_mode"); }
  pl->kernel = split ? (ConvGemmKernel)his_gemm_pick_split_kernel(bk, actc, res_mode, EPI_PLAIN, halo) : pick_kernel(bk, actc, res_mode, halo);
  pl->smem = smem_for(bk);
  p.in = (const __half*)in; p.in_sn = (long long)H * W * in_cs; p.in_cs = in_cs; p.cin = cin;
  p.a_stages = 0; p.a_stage_bytes = 0; p.taps_per_b = 1; p.taps_per_box = 1;
  p.out = (__half*)out; p.out_cs = out_cs; p.res = (const __half*)res; p.res_cs = res_cs; p.direct_ok = transposed ? 0 : 1;
  if (const char* e = getenv("HIS_GEMM_DIRECT")) p.direct_ok = p.direct_ok ? atoi(e) : 0;   // 0 never, 1 channel-clipped chunks, 2 + clipped tiles, 3 = 1 + residual of clipped tiles
  p.debug = 0;
  if (const char* e = getenv("HIS_GEMM_DEBUG")) p.debug = atoi(e);
  p.dbg_ts = nullptr;
  if (p.debug & 16) {
    if (cudaMalloc(&p.dbg_ts, 64 * 32 * sizeof(unsigned long long)) == cudaSuccess) cudaMemset(p.dbg_ts, 0, 64 * 32 * sizeof(unsigned long long));
    else p.dbg_ts = nullptr;
  }
  p.up_in = (const __half*)in; p.up_sn = 0; p.up_cs = 8; p.up_split = 0;      // no fused upsample
  // TMEM accumulator ring: as many buffers as the 512 columns hold (even, <= 8) so that short tiles are not paced by
  // the MMA -> epilogue -> MMA hand-shake latency
  p.acc_stride = (p.block_n + 31) / 32 * 32;
  p.n_acc = kTmemCols / p.acc_stride;
  if (p.n_acc > 8) p.n_acc = 8;
  p.n_acc &= ~1;
  if (const char* e = getenv("HIS_GEMM_NACC")) { int v = atoi(e) & ~1; if (v >= 2 && v <= p.n_acc) p.n_acc = v; }
  p.fast_decode = (p.n_tiles == 1 && p.groups == 1 && p.num_work < (1 << 21)) ? 1 : 0;
  // a third epilogue warpgroup for the layers bound by the epilogue's per-tile latency: narrow N tiles and the 1x1 convs
  // (HIS_GEMM_EPI=2|3 forces the count; pair kernels keep two)
  p.epi_groups = (p.block_n <= 96 || (ksize == 1 && cin <= 112) || (k1_split && ksize == 1 && !split && p.block_n == 128 && (p.n_tiles > 1 || transposed))) ? 3 : 2;
  if (const char* e = getenv("HIS_GEMM_EPI")) { const int v = atoi(e); if (v == 2 || v == 3) p.epi_groups = v; }
  // every accumulator must always be drained by the SAME group (a group that met an accumulator first at its second use would
  // pass the parity wait of the still untouched barrier): the ring length is a multiple of the group count
  if (p.epi_groups == 3) {
    const int fit = kTmemCols / p.acc_stride;
    if (fit >= 6) p.n_acc = 6; else if (fit >= 3) p.n_acc = 3; else p.epi_groups = 2;
  }
  p.inv_tiles_x = 1.0f / (float)p.tiles_x; p.inv_tiles_y = 1.0f / (float)p.tiles_y;
  if (halo) {
    // B ring stage = taps_per_b weight tiles (block_n x BK); A ring stage = BK/8 planes of the 10 x 18 window
    // every barrier round trip of a B stage costs the issuing thread ~300 ns, so a stage carries as many taps as 48 KB hold
    p.pair = pair_halo ? 1 : 0;
    const int tap_bytes = (p.pair ? p.block_n / 2 : p.block_n) * bk * 2;      // the weight rows THIS CTA stages per tap
    p.taps_per_b = 9 * tap_bytes <= 48 * 1024 ? 9 : 3 * tap_bytes <= 48 * 1024 ? 3 : 1;
    if (const char* e = getenv("HIS_GEMM_TAPS")) { int v = atoi(e); if (v == 1 || v == 3 || v == 9) p.taps_per_b = v; }
    p.taps_per_box = (p.n_tiles == 1 && p.taps_per_b * p.block_n <= 256) ? p.taps_per_b : 1;
    if (p.pair) p.taps_per_box = 1;      // a CTA's half of a tap's rows is not contiguous with the next tap's
    p.stage_bytes = (p.taps_per_b * tap_bytes + 1023) / 1024 * 1024;
    p.a_stage_bytes = ((bk / 8) * kPlaneBytes + 127) / 128 * 128;   // stage pitch: 128-byte aligned (TMA destination); the TMA window
                                                                    // uses 16 B less per plane of it
    const int budget = KCfg<64>::kRingBytes - (p.epi_groups - 2) * 2 * kStagingBytes;
    int a_st = p.a_stage_bytes >= 16 * 1024 ? (p.pair ? 3 : 2) : (budget / 3) / p.a_stage_bytes;
    if (a_st > kMaxAStages) a_st = kMaxAStages;
    if (a_st < 2) a_st = 2;
    if (const char* e = getenv("HIS_GEMM_ASTAGES")) { int v = atoi(e); if (v >= 1 && v <= kMaxAStages) a_st = v; }
    int b_st = (budget - a_st * p.a_stage_bytes) / p.stage_bytes;
    if (b_st > kMaxBStages) b_st = kMaxBStages;
    if (b_st < 2) { delete pl; return his_set_error(HIS_ERR_UNSUPPORTED, "conv_gemm halo mode: shared memory budget"); }
    p.a_stages = a_st; p.stages = b_st;
    // weight-resident mode: the whole [K block][tap] weight set fits next to a deep halo ring -> loaded once per CTA
    p.b_resident = 0;
    const int res_bytes = (p.kblocks_per_tap * 9 * p.block_n * bk * 2 + 1023) / 1024 * 1024;
    int want_res = p.n_tiles == 1 && res_bytes <= 96 * 1024 && !p.pair;
    if (const char* e = getenv("HIS_GEMM_BRES")) want_res = want_res && atoi(e) != 0;
    if (want_res) {
      int a2 = (budget - res_bytes) / p.a_stage_bytes;
      if (a2 > kMaxAStages) a2 = kMaxAStages;
      if (a2 >= 2) { p.b_resident = 1; p.stages = 1; p.stage_bytes = res_bytes; p.a_stages = a2; p.taps_per_b = 9; }
    }
  } else {
    // CTA pairs (cta_group::2): two adjacent pixel tiles of one image share the weight tile, each CTA stages half of it -> half
    // the shared-memory operand traffic per MMA for B and half the weight L2 traffic.  Needs a single N tile (adjacent work
    // items = adjacent pixel tiles), an even number of tiles per image and block_n a multiple of 32.
    int pair_min_n = 128;
    if (const char* e = getenv("HIS_GEMM_PAIR")) pair_min_n = atoi(e) > 0 ? atoi(e) : (1 << 30);
    // (1x1 layers are HBM / L2 bound, the pair's lock step only costs there: measured 0.50 -> 0.55 ms on 256->256 k1)
    p.pair = (!split && !transposed && ksize == 3 && p.n_tiles == 1 && bk == 64 && p.block_n >= pair_min_n && (p.block_n % 32) == 0 &&
              ((p.tiles_x * p.tiles_y) % 2) == 0 && p.num_work >= 2) ? 1 : 0;
    p.stage_bytes = (kBlockM * bk * 2 + (p.pair ? p.block_n / 2 : p.block_n) * bk * 2 + 1023) / 1024 * 1024;
    p.stages = (KCfg<64>::kRingBytes - (p.epi_groups - 2) * 2 * kStagingBytes) / p.stage_bytes;
    if (p.stages > kMaxStages) p.stages = kMaxStages;
    if (const char* e = getenv("HIS_GEMM_STAGES")) { int v = atoi(e); if (v >= 2 && v < p.stages) p.stages = v; }
  }
  pl->actc = actc; pl->res_mode = res_mode; pl->transposed = transposed; pl->cin_pad = cin_pad;
  p.tail_c = 0; p.store_main = 1; p.aux_out = nullptr; p.aux_bufs = 1; p.cout = cout;
  p.shift = shift;
  int taps = ksize * ksize;
  int rc;
  if ((rc = encode_act_map(&pl->tmA, in, cin, W, H, n_img, in_cs, (long long)W * in_cs, (long long)H * W * in_cs, bk, p.bw, p.bh, p.in_lo))) { delete pl; return rc; }
  // halo window by TMA: one 5-D box {8, 10, 18, BK/8, 1} of the map {8 ch, W, H, C/8, N}.  HIS_GEMM_HALO_TMA=0 keeps the cp.async producers.
  p.a_tma = 0; p.plane_bytes = kPlaneBytes;
  if (halo && !split && !p.pair) {
    int want = 1;
    if (const char* e = getenv("HIS_GEMM_HALO_TMA")) want = atoi(e) != 0;
    PFN_encodeTiled enc = get_encode();
    if (want && enc) {
      cuuint64_t dims[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(cin / 8), (cuuint64_t)n_img};
      cuuint64_t strides[4] = {(cuuint64_t)in_cs * 2, (cuuint64_t)W * in_cs * 2, 16, (cuuint64_t)H * W * in_cs * 2};
      cuuint32_t box[5] = {8, (cuuint32_t)kHaloW, (cuuint32_t)kHaloH, (cuuint32_t)(bk / 8), 1};
      cuuint32_t es[5] = {1, 1, 1, 1, 1};
      if (enc(&pl->tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(in), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
        p.a_tma = 1; p.plane_bytes = kHaloPix * 16;
      }
    }
  }
  pl->b_box_rows = p.pair ? p.block_n / 2 : halo ? p.taps_per_box * p.block_n : p.block_n;
  if ((rc = encode_weight_map(&pl->tmB, w_packed, cin_pad, (long long)p.groups * taps * p.cout_slab, pl->b_box_rows, bk))) { delete pl; return rc; }
  if (!transposed) {
    if ((rc = encode_act_map(&pl->tmO[0], out, cout, W, H, n_img, out_cs, (long long)W * out_cs, (long long)H * W * out_cs, kChunkC, p.bw, p.bh, p.out_lo))) { delete pl; return rc; }
    pl->tmO[1] = pl->tmO[2] = pl->tmO[3] = pl->tmO[0];
  } else {
    for (int g = 0; g < 4; ++g) {
      int dy = g >> 1, dx = g & 1;
      const __half* base = (const __half*)out + ((long long)dy * 2 * W + dx) * out_cs;
      if ((rc = encode_act_map(&pl->tmO[g], base, cout, W, H, n_img, 2LL * out_cs, 4LL * W * out_cs, 4LL * H * W * out_cs, kChunkC, p.bw, p.bh, p.out_lo))) { delete pl; return rc; }
    }
  }
  if (res_mode != HIS_RES_NONE) {
    if ((rc = encode_act_map(&pl->tmR, res, cout, W, H, n_img, res_cs, (long long)W * res_cs, (long long)H * W * res_cs, kChunkC, p.bw, p.bh, p.res_lo))) { delete pl; return rc; }
  } else {
    pl->tmR = pl->tmO[0];
  }
  pl->tmAux = pl->tmO[0];
  pl->grid = p.num_work < num_sms ? p.num_work : num_sms;
  if (p.pair) { pl->grid &= ~1; pl->kernel = pick_pair_kernel(actc, res_mode, halo); }
  *out_plan = pl;
  return HIS_OK;
}

// The fused tail and the per-pixel statistics need the whole channel range in ONE N tile: a 1x1 layer that was split into two
// 128-wide tiles goes back to its 256-wide tile (two epilogue groups, accumulator ring of two) -- before anything that depends on
// the work-item count or the ring depth was attached to the plan.
static int unsplit_k1(ConvGemmPlan* pl, const char* who) {
  if (!pl->k1_split) return HIS_OK;
  ConvGemmParams& p = pl->p;
  if (p.aux_out || p.ln_partials) return his_set_error(HIS_ERR_UNSUPPORTED, who);
  p.n_tiles = 1; p.block_n = 256; p.cout_slab = 256; p.phase_slab = 256;
  p.num_work = p.n_img * p.tiles_y * p.tiles_x;
  p.acc_stride = 256; p.n_acc = 2; p.epi_groups = 2;
  p.fast_decode = p.num_work < (1 << 21) ? 1 : 0;
  p.stage_bytes = (kBlockM * pl->bk * 2 + 256 * pl->bk * 2 + 1023) / 1024 * 1024;
  p.stages = KCfg<64>::kRingBytes / p.stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  pl->b_box_rows = 256;
  pl->grid = p.num_work < pl->num_sms ? p.num_work : pl->num_sms;
  pl->k1_split = 0;
  return encode_weight_map(&pl->tmB, pl->w_packed, pl->cin_pad, (long long)p.cout_slab, pl->b_box_rows, pl->bk);
}

int his_conv_gemm_set_tail(void* plan, const float* tail_w, float tail_b0, float tail_b1, int tail_c, int tail_sigmoid, float* tail_out,
                           int store_main) {
  if (!plan || !tail_w || !tail_out) return his_set_error(HIS_ERR_INVALID_ARG, "set_tail: null pointer");
  ConvGemmPlan* pl = (ConvGemmPlan*)plan;
  if (tail_c < 1 || tail_c > 2) return his_set_error(HIS_ERR_INVALID_ARG, "set_tail: tail_c must be 1 or 2");
  if (int rc = unsplit_k1(pl, "set_tail: attach the tail before the fp32 export / LayerNorm statistics of a 1x1 layer")) return rc;
  if (pl->p.n_tiles != 1 || pl->transposed || pl->actc != ACTC_CLAMP || pl->res_mode == HIS_RES_MUL)
    return his_set_error(HIS_ERR_UNSUPPORTED, "set_tail: needs a single N tile, none/relu activation, no MUL operand, not transposed");
  pl->p.tail_w = tail_w; pl->p.tail_b0 = tail_b0; pl->p.tail_b1 = tail_b1; pl->p.tail_c = tail_c; pl->p.tail_sigmoid = tail_sigmoid;
  pl->p.tail_out = tail_out; pl->p.store_main = store_main;
  pl->kernel = pl->split ? (ConvGemmKernel)his_gemm_pick_split_kernel(pl->bk, ACTC_CLAMP, pl->res_mode, EPI_TAIL, pl->halo)
               : pl->p.pair ? pick_pair_tail_kernel(pl->res_mode, pl->halo) : pick_tail_kernel(pl->bk, pl->res_mode, pl->halo);
  return HIS_OK;
}

int his_conv_gemm_set_image_weights(void* plan, const void* w_packed_per_image) {
  if (!plan || !w_packed_per_image) return his_set_error(HIS_ERR_INVALID_ARG, "set_image_weights: null pointer");
  ConvGemmPlan* pl = (ConvGemmPlan*)plan;
  ConvGemmParams& p = pl->p;
  if (pl->halo && p.b_resident) return his_set_error(HIS_ERR_UNSUPPORTED, "set_image_weights: not for weight-resident halo layers");
  const long long rows = (long long)p.groups * p.ksize * p.ksize * p.cout_slab;
  if (rows * p.n_img >= (1LL << 31)) return his_set_error(HIS_ERR_UNSUPPORTED, "set_image_weights: too many weight rows");
  int rc = encode_weight_map(&pl->tmB, w_packed_per_image, pl->cin_pad, rows * p.n_img, pl->b_box_rows, pl->bk);
  if (rc) return rc;
  p.b_img_rows = (int)rows;
  return HIS_OK;
}

int his_conv_gemm_set_res_scale(void* plan, const float* res_scale) {
  if (!plan) return his_set_error(HIS_ERR_INVALID_ARG, "set_res_scale: null plan");
  ConvGemmPlan* pl = (ConvGemmPlan*)plan;
  if (res_scale && pl->res_mode == HIS_RES_NONE) return his_set_error(HIS_ERR_INVALID_ARG, "set_res_scale: the layer has no residual operand");
  pl->p.res_scale = res_scale;
  return HIS_OK;
}

int his_conv_gemm_set_row_ops(void* plan, const float* row_scale, float* stats_out) {
  if (!plan) return his_set_error(HIS_ERR_INVALID_ARG, "set_row_ops: null plan");
  ConvGemmPlan* pl = (ConvGemmPlan*)plan;
  if (stats_out)
    if (int rc = unsplit_k1(pl, "set_row_ops: attach the statistics before the fp32 export / LayerNorm statistics of a 1x1 layer")) return rc;
  if (stats_out && (pl->p.n_tiles != 1 || pl->transposed)) return his_set_error(HIS_ERR_UNSUPPORTED, "set_row_ops: statistics need a single N tile, not transposed");
  pl->p.row_scale = row_scale; pl->p.stats_out = stats_out;
  return HIS_OK;
}

int his_conv_gemm_set_upsampled_input(void* plan, const void* low, int low_c, int low_cs) {
  if (!plan || !low) return his_set_error(HIS_ERR_INVALID_ARG, "set_upsampled_input: null pointer");
  ConvGemmPlan* pl = (ConvGemmPlan*)plan;
  ConvGemmParams& p = pl->p;
  if (!pl->halo || (p.H & 1) || (p.W & 1) || low_c <= 0 || low_c > p.cin || (low_c % pl->bk) || (low_cs % 8) || low_c > low_cs)
    return his_set_error(HIS_ERR_UNSUPPORTED, "set_upsampled_input: needs a halo-mode 3x3 layer, even H and W, low_c a multiple of the K block");
  p.up_in = (const __half*)low; p.up_cs = low_cs; p.up_sn = (long long)(p.H >> 1) * (p.W >> 1) * low_cs; p.up_split = low_c;
  p.up_lo = pl->split ? low_cs / 2 : 0;
  p.a_tma = 0; p.plane_bytes = kPlaneBytes;      // the gather from two tensors needs the cp.async producers
  return HIS_OK;
}

// 1 when his_conv_gemm_set_upsampled_input would accept this layer (same policy as his_conv_gemm_create)
int his_conv_gemm_can_fuse_upsample(int H, int W, int cin, int cout, int low_c) {
  int nt = 0, bn = 0;
  if (his_conv_gemm_tile_n(cout, &nt, &bn) != HIS_OK) return 0;
  const int halo_maxn = halo_max_n();
  if (const char* e = getenv("HIS_GEMM_HALO")) if (atoi(e) == 0) return 0;
  if (const char* e = getenv("HIS_GEMM_FUSE_UP")) if (atoi(e) == 0) return 0;
  if ((cin % 8) || bn > halo_maxn || (H & 1) || (W & 1)) return 0;
  const int bk = cin > 32 ? 64 : cin > 16 ? 32 : 16;
  return (low_c > 0 && low_c <= cin && low_c % bk == 0) ? 1 : 0;
}

int his_conv_gemm_set_ln_partials(void* plan, double* partials, int* parts_per_image) {
  if (!plan || !partials || !parts_per_image) return his_set_error(HIS_ERR_INVALID_ARG, "set_ln_partials: null pointer");
  ConvGemmPlan* pl = (ConvGemmPlan*)plan;
  ConvGemmParams& p = pl->p;
  if (p.n_img <= 0 || p.num_work % p.n_img) return his_set_error(HIS_ERR_UNSUPPORTED, "set_ln_partials: work items do not split by image");
  if (p.tail_c && !p.store_main) return his_set_error(HIS_ERR_UNSUPPORTED, "set_ln_partials: the layer does not store its output");
  if (p.res_scale) return his_set_error(HIS_ERR_UNSUPPORTED, "set_ln_partials: not together with a residual scale");
  p.ln_partials = partials;
  *parts_per_image = p.num_work / p.n_img;
  return HIS_OK;
}

int his_conv_gemm_work_items(void* plan) { return plan ? ((ConvGemmPlan*)plan)->p.num_work : 0; }

int his_conv_gemm_set_aux(void* plan, float* aux_out) {
  if (!plan || !aux_out) return his_set_error(HIS_ERR_INVALID_ARG, "set_aux: null pointer");
  ConvGemmPlan* pl = (ConvGemmPlan*)plan;
  if (pl->bk != 64 || pl->transposed || pl->p.tail_c)
    return his_set_error(HIS_ERR_UNSUPPORTED, "set_aux: needs a 64-wide K block (Cin >= 64), not transposed, no fused tail");
  pl->p.aux_out = aux_out;
  // export through a [32 ch][128 px] fp32 staging tile + one TMA store per chunk when the rows are 16-byte addressable and the
  // 32 KB of staging can be taken from the operand rings (HIS_GEMM_AUX_TMA=0: per-thread strided stores)
  ConvGemmParams& p = pl->p;
  int want = (p.W % 4) == 0;
  if (const char* e = getenv("HIS_GEMM_AUX_TMA")) want = want && atoi(e) != 0;
  p.aux_tma = 0; p.aux_bufs = 1;
  if (want) {
    // two export tiles per group (the store of a chunk drains while the next chunk is written) where the operand ring can spare
    // the 16 KB per group: the short K loops of the 1x1 layers; a 3x3 layer keeps its ring depth and one tile
    bool ok = false;
    int max_bufs = p.ksize == 1 ? 2 : 1;
    if (const char* e = getenv("HIS_GEMM_AUX_BUFS")) { const int v = atoi(e); if (v == 1 || v == 2) max_bufs = v; }
    for (int bufs = max_bufs; bufs >= 1 && !ok; --bufs) {
      const int need = p.epi_groups * bufs * kAuxStagingBytes;
      if (!pl->halo) {
        const int st = (KCfg<64>::kRingBytes - (p.epi_groups - 2) * 2 * kStagingBytes - need) / p.stage_bytes;
        if (st >= 2) { if (st < p.stages) p.stages = st; ok = true; }
      } else {
        const int budget = KCfg<64>::kRingBytes - (p.epi_groups - 2) * 2 * kStagingBytes - need;
        int a_st = p.a_stages, b_st = p.stages;
        while (b_st * p.stage_bytes + a_st * p.a_stage_bytes > budget && a_st > 2) --a_st;
        while (b_st * p.stage_bytes + a_st * p.a_stage_bytes > budget && b_st > 2 && !p.b_resident) --b_st;
        if (b_st * p.stage_bytes + a_st * p.a_stage_bytes <= budget) { p.a_stages = a_st; p.stages = b_st; ok = true; }
      }
      if (ok) p.aux_bufs = bufs;
    }
    if (ok) {
      PFN_encodeTiled enc = get_encode();
      cuuint64_t dims[4] = {(cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.cout, (cuuint64_t)p.n_img};
      cuuint64_t strides[3] = {(cuuint64_t)p.W * 4, (cuuint64_t)p.H * p.W * 4, (cuuint64_t)p.cout * p.H * p.W * 4};
      cuuint32_t box[4] = {(cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)kChunkC, 1};
      cuuint32_t es[4] = {1, 1, 1, 1};
      if (enc && !((uintptr_t)aux_out & 15) &&
          enc(&pl->tmAux, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, aux_out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
        p.aux_tma = 1;
    }
  }
  pl->kernel = pl->split ? (ConvGemmKernel)his_gemm_pick_split_kernel(64, pl->actc, pl->res_mode, EPI_AUX, pl->halo)
               : pl->p.pair ? pick_pair_aux_kernel(pl->actc, pl->res_mode, pl->halo) : pick_aux_kernel(pl->actc, pl->res_mode, pl->halo);
  return HIS_OK;
}

int his_conv_gemm_run(void* plan, void* stream) {
  if (!plan) return his_set_error(HIS_ERR_INVALID_ARG, "null plan");
  ConvGemmPlan* pl = (ConvGemmPlan*)plan;
  if (pl->p.num_work == 0) return HIS_OK;
  if (pl->p.pair) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(pl->grid); cfg.blockDim = dim3(128 + 128 * pl->p.epi_groups); cfg.dynamicSmemBytes = pl->smem; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    if (!pl->pair_grid_checked) {      // persistent kernel: never more clusters than can be co-resident (a TPC with one SM hosts none)
      int ncl = 0;
      if (cudaOccupancyMaxActiveClusters(&ncl, pl->kernel, &cfg) == cudaSuccess && ncl > 0 && 2 * ncl < pl->grid) pl->grid = 2 * ncl;
      (void)cudaGetLastError();
      cfg.gridDim = dim3(pl->grid);
      pl->pair_grid_checked = 1;
    }
    if (cudaLaunchKernelEx(&cfg, pl->kernel, pl->tmA, pl->tmB, pl->tmO[0], pl->tmO[1], pl->tmO[2], pl->tmO[3], pl->tmR, pl->tmAux, pl->p) != cudaSuccess)
      return his_set_error(HIS_ERR_LAUNCH, cudaGetErrorString(cudaGetLastError()));
    return HIS_OK;
  }
  pl->kernel<<<pl->grid, 128 + 128 * pl->p.epi_groups, pl->smem, (cudaStream_t)stream>>>(pl->tmA, pl->tmB, pl->tmO[0], pl->tmO[1], pl->tmO[2], pl->tmO[3], pl->tmR,
                                                               pl->tmAux, pl->p);
  HIS_CHECK_LAUNCH();
  return HIS_OK;
}

int his_conv_gemm_destroy(void* plan) {
  ConvGemmPlan* pl = (ConvGemmPlan*)plan;
  if (pl && pl->p.dbg_ts) {
    // tuning aid (HIS_GEMM_DEBUG & 16): mean clock deltas between the epilogue's stamps over the tiles thread 128 of CTA 0 drained
    static unsigned long long ts[64 * 32];
    cudaDeviceSynchronize();
    if (cudaMemcpy(ts, pl->p.dbg_ts, sizeof ts, cudaMemcpyDeviceToHost) == cudaSuccess) {
      int nt = 0;
      while (nt < 64 && ts[nt * 32 + 24]) ++nt;
      fprintf(stderr, "[his dbg] N%d %dx%d cin%d cout%d k%d res%d tiles stamped %d", pl->p.n_img, pl->p.H, pl->p.W, pl->p.cin, pl->p.cout, pl->p.ksize,
              pl->res_mode, nt);
      if (nt > 4) {
        fprintf(stderr, " period %.0f clk |", (double)(ts[(nt - 1) * 32] - ts[2 * 32]) / (nt - 3));
        int prev = 0;
        for (int k = 1; k <= 24; ++k) {
          double sum = 0; int cnt = 0;
          for (int t = 2; t < nt; ++t)
            if (ts[t * 32 + k] && ts[t * 32 + prev]) { sum += (double)(long long)(ts[t * 32 + k] - ts[t * 32 + prev]); ++cnt; }
          if (cnt) { fprintf(stderr, " %d:%.0f", k, sum / cnt); prev = k; }
        }
      }
      fprintf(stderr, "\n");
    }
    cudaFree(pl->p.dbg_ts);
  }
  delete pl;
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
