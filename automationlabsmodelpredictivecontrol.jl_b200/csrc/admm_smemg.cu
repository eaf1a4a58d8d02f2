// Launchers of the shared-memory resident general-row ADMM kernel (admm_smemg.cuh) in their own translation unit.
#include <algorithm>

#include "admm_smemg.cuh"

namespace mpcb {

// the smallest configuration (2 warps) decides whether a controller fits at all
size_t smemg_bytes_host(int NT, int np, bool sig) { return smemg_bytes(NT, np, sig, 3); }

namespace {
template <int NT, bool SIG, int W>
cudaError_t launch_smemg_w(const OnchipParams& P, int sm_count, cudaStream_t st) {
  auto kern = mpcb::admm_smemg_kernel<NT, SIG, W>;
  const size_t smem = mpcb::smemg_bytes(NT, P.np, SIG, W);
  static bool attr_set[64] = {};      // function attributes are per device (a multi-device handle launches on several)
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  const long long blocks_needed = (P.batch + 8 * W - 1) / (8 * W);
  const long long grid = std::min<long long>(blocks_needed, (long long)sm_count);      // one CTA per SM (shared-memory bound)
  kern<<<(unsigned)std::max<long long>(grid, 1), W * 32, smem, st>>>(P);
  return cudaGetLastError();
}

template <int NT, bool SIG>
cudaError_t launch_smemg_t(const OnchipParams& P, int sm_count, cudaStream_t st) {
  // as many warps as T and their state slices fit in (227 KB = 232448 B of opt-in shared memory per CTA)
  if (mpcb::smemg_bytes(NT, P.np, SIG, 8) <= 232448) return launch_smemg_w<NT, SIG, 8>(P, sm_count, st);
  if (mpcb::smemg_bytes(NT, P.np, SIG, 6) <= 232448) return launch_smemg_w<NT, SIG, 6>(P, sm_count, st);
  if (mpcb::smemg_bytes(NT, P.np, SIG, 5) <= 232448) return launch_smemg_w<NT, SIG, 5>(P, sm_count, st);
  if (mpcb::smemg_bytes(NT, P.np, SIG, 4) <= 232448) return launch_smemg_w<NT, SIG, 4>(P, sm_count, st);
  return launch_smemg_w<NT, SIG, 3>(P, sm_count, st);
}
}  // namespace

cudaError_t launch_smemg(int NT, const OnchipParams& P, int sm_count, cudaStream_t st) {
  const bool sig = P.sigma != 0.0;
  switch (NT) {
#define MPCB_SG(N_) case N_: return sig ? launch_smemg_t<N_, true>(P, sm_count, st) : launch_smemg_t<N_, false>(P, sm_count, st);
    MPCB_SG(32) MPCB_SG(40) MPCB_SG(48) MPCB_SG(56) MPCB_SG(64) MPCB_SG(72) MPCB_SG(80) MPCB_SG(88) MPCB_SG(96) MPCB_SG(104) MPCB_SG(112) MPCB_SG(120)
#undef MPCB_SG
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace mpcb
