// Launcher of the CTA-cooperative straggler kernel (admm_coop.cuh) in its own translation unit.
#include <algorithm>

#include "admm_coop.cuh"

namespace mpcb {

size_t coop_bytes_host(int NT, int np, bool sig) { return coop_bytes(NT, np, sig); }

namespace {
template <int NT, bool SIG>
cudaError_t launch_coop_t(const OnchipParams& P, int sm_count, cudaStream_t st) {
  auto kern = mpcb::admm_coop_kernel<NT, SIG>;
  const size_t smem = mpcb::coop_bytes(NT, P.np, SIG);
  static bool attr_set[64] = {};      // function attributes are per device
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  // one CTA per SM at most (the ticket count lives on the device: surplus CTAs find the queue empty and exit)
  const long long groups = P.tickets_max >= 0 ? (P.tickets_max + 7) / 8 : (P.batch + 7) / 8;
  const long long grid = std::max<long long>(1, std::min<long long>(groups, (long long)sm_count));
  kern<<<(unsigned)grid, COOP_WARPS * 32, smem, st>>>(P);
  return cudaGetLastError();
}
}  // namespace

cudaError_t launch_coop(int NT, const OnchipParams& P, int sm_count, cudaStream_t st) {
  const bool sig = P.sigma != 0.0;
  switch (NT) {
#define MPCB_CO(N_) case N_: return sig ? launch_coop_t<N_, true>(P, sm_count, st) : launch_coop_t<N_, false>(P, sm_count, st);
    MPCB_CO(24) MPCB_CO(32) MPCB_CO(40) MPCB_CO(48) MPCB_CO(56) MPCB_CO(64) MPCB_CO(72) MPCB_CO(80) MPCB_CO(88) MPCB_CO(96) MPCB_CO(104) MPCB_CO(112) MPCB_CO(120)
#undef MPCB_CO
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace mpcb
