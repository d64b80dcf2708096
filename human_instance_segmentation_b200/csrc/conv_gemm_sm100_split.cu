// Split-fp16 ("strict" precision) instantiations of the implicit-GEMM convolution: the same kernel template as
// conv_gemm_sm100.cu with SPLIT = true (activations as fp16 hi + lo planes, weights [W_hi | W_lo], three MMA passes per K block,
// fp32 accumulation in TMEM, hi / lo re-split in the epilogue).  Matches the reference's fp32 evaluation
// (hed/train_utils.py:162-180, no autocast) to ~1e-5 instead of the ~1.5e-3 of single fp16 operands.  A separate translation
// unit so the two sets of instantiations compile in parallel.
#include "conv_gemm_sm100_kernel.cuh"

namespace {

template <int BK, int ACTC, int RES, int EPI, bool HALO>
ConvGemmKernel split_k() { return conv_gemm_sm100_kernel<BK, ACTC, RES, EPI, HALO, false, true>; }

template <int BK, int ACTC, int EPI, bool HALO>
ConvGemmKernel split_res(int res) {
  return res == 0 ? split_k<BK, ACTC, 0, EPI, HALO>() : res == 1 ? split_k<BK, ACTC, 1, EPI, HALO>() : split_k<BK, ACTC, 2, EPI, HALO>();
}
template <int BK, int EPI, bool HALO>
ConvGemmKernel split_act(int actc, int res) {
  return actc == 0 ? split_res<BK, 0, EPI, HALO>(res) : actc == 1 ? split_res<BK, 1, EPI, HALO>(res) : split_res<BK, 2, EPI, HALO>(res);
}
template <int EPI, bool HALO>
ConvGemmKernel split_bk(int bk, int actc, int res) {
  return bk == 64 ? split_act<64, EPI, HALO>(actc, res) : bk == 32 ? split_act<32, EPI, HALO>(actc, res) : split_act<16, EPI, HALO>(actc, res);
}
// fused-tail variants: clamp activations, no residual or residual-add (like the single-fp16 set)
template <int BK, bool HALO>
ConvGemmKernel split_tail(int res) { return res == 0 ? split_k<BK, ACTC_CLAMP, 0, EPI_TAIL, HALO>() : split_k<BK, ACTC_CLAMP, 1, EPI_TAIL, HALO>(); }
template <bool HALO>
ConvGemmKernel split_tail_bk(int bk, int res) {
  return bk == 64 ? split_tail<64, HALO>(res) : bk == 32 ? split_tail<32, HALO>(res) : split_tail<16, HALO>(res);
}

ConvGemmKernel pick(int bk, int actc, int res, int epi, int halo) {
  if (epi == EPI_TAIL) return (res > 1 || actc != ACTC_CLAMP) ? nullptr : halo ? split_tail_bk<true>(bk, res) : split_tail_bk<false>(bk, res);
  if (epi == EPI_AUX) return bk != 64 ? nullptr : halo ? split_act<64, EPI_AUX, true>(actc, res) : split_act<64, EPI_AUX, false>(actc, res);
  return halo ? split_bk<EPI_PLAIN, true>(bk, actc, res) : split_bk<EPI_PLAIN, false>(bk, actc, res);
}

}  // namespace

extern "C" void* his_gemm_pick_split_kernel(int bk, int actc, int res, int epi, int halo) { return (void*)pick(bk, actc, res, epi, halo); }

extern "C" int his_gemm_split_set_smem_attr(void) {
  for (int h = 0; h < 2; ++h)
    for (int epi = 0; epi < 3; ++epi)
      for (int bk = 16; bk <= 64; bk *= 2)
        for (int a = 0; a < 3; ++a)
          for (int r = 0; r < 3; ++r) {
            ConvGemmKernel k = pick(bk, a, r, epi, h);
            if (k && cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget) != cudaSuccess)
              return his_set_error(HIS_ERR_LAUNCH, "cannot raise dynamic shared memory limit (split kernels)");
          }
  return HIS_OK;
}
