// Launchers of the shared-memory resident ADMM kernel (admm_smem.cuh) in their own translation unit.
#include <algorithm>

#include "admm_smem.cuh"

namespace mpcb {

size_t smemk_bytes_host(int NT, int np, bool sig) { return smemk_bytes(NT, np, sig, 4); }

namespace {
template <int NT, bool SIG, int W>
cudaError_t launch_smemk_w(const OnchipParams& P, int sm_count, int* attr_set, cudaStream_t st) {
  auto kern = mpcb::admm_smem_kernel<NT, SIG, W>;
  const size_t smem = mpcb::smemk_bytes(NT, P.np, SIG, W);
  if (*attr_set == 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);      // the device maximum: another handle with a longer parameter vector must not find the cap lowered
    if (e != cudaSuccess) return e;
    *attr_set = 1;
  }
  const long long blocks_needed = (P.batch + 8 * W - 1) / (8 * W);
  const long long grid = std::min<long long>(blocks_needed, (long long)sm_count);      // one CTA per SM (shared-memory bound)
  kern<<<(unsigned)std::max<long long>(grid, 1), W * 32, smem, st>>>(P);
  return cudaGetLastError();
}

template <int NT, bool SIG>
cudaError_t launch_smemk_t(const OnchipParams& P, int sm_count, int* attr_set, cudaStream_t st) {
  // as many warps (8, 6, 5 or 4) as T and their state slices fit in (227 KB = 232448 B of opt-in shared memory per CTA)
  if (mpcb::smemk_bytes(NT, P.np, SIG, 8) <= 232448) return launch_smemk_w<NT, SIG, 8>(P, sm_count, attr_set, st);
  if (mpcb::smemk_bytes(NT, P.np, SIG, 6) <= 232448) return launch_smemk_w<NT, SIG, 6>(P, sm_count, attr_set, st);
  if (mpcb::smemk_bytes(NT, P.np, SIG, 5) <= 232448) return launch_smemk_w<NT, SIG, 5>(P, sm_count, attr_set, st);
  return launch_smemk_w<NT, SIG, 4>(P, sm_count, attr_set, st);
}

}  // namespace

cudaError_t launch_smemk(int NT, const OnchipParams& P, int sm_count, int* attr_set, cudaStream_t st) {
  const bool sig = P.sigma != 0.0;
  switch (NT) {
#define MPCB_SK(N_) case N_: return sig ? launch_smemk_t<N_, true>(P, sm_count, attr_set, st) : launch_smemk_t<N_, false>(P, sm_count, attr_set, st);
    MPCB_SK(72) MPCB_SK(80) MPCB_SK(88) MPCB_SK(96) MPCB_SK(104) MPCB_SK(112) MPCB_SK(120)
#undef MPCB_SK
    default: return cudaErrorInvalidValue;
  }
}


}  // namespace mpcb
