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
  static size_t attr_bytes[64] = {};      // dynamic shared memory opted into so far, per device (the size depends on the parameter length np)
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (attr_bytes[dev] == 0) {      // opt into the device maximum once: the size a launch needs depends on np, and lowering the cap later would break a larger controller
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return e;
    attr_bytes[dev] = 232448;
  }
  // one CTA per SM at most (the ticket count lives on the device: surplus CTAs find the queue empty and exit)
  const long long groups = P.tickets_max >= 0 ? (P.tickets_max + 7) / 8 : (P.batch + 7) / 8;
  const long long grid = std::max<long long>(1, std::min<long long>(groups, (long long)sm_count));
  kern<<<(unsigned)grid, COOP_WARPS * 32, smem, st>>>(P);
  return cudaGetLastError();
}
template <int NT, bool SIG>
cudaError_t launch_coopb_t(const OnchipParams& P, int sm_count, cudaStream_t st) {
  auto kern = mpcb::admm_coopb_kernel<NT, SIG>;
  const size_t smem = mpcb::coopb_bytes(NT, P.np, SIG);
  static size_t attr_bytes[64] = {};      // dynamic shared memory opted into so far, per device (the size depends on the parameter length np)
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (attr_bytes[dev] == 0) {      // opt into the device maximum once: the size a launch needs depends on np, and lowering the cap later would break a larger controller
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return e;
    attr_bytes[dev] = 232448;
  }
  const long long grid = std::max<long long>(1, std::min<long long>((P.batch + 7) / 8, (long long)sm_count));
  kern<<<(unsigned)grid, COOP_WARPS * 32, smem, st>>>(P);
  return cudaGetLastError();
}
}  // namespace

size_t coopb_bytes_host(int NT, int np, bool sig) { return coopb_bytes(NT, np, sig); }

// box-only small batches (no remap / device-side ticket count: the host knows the batch)
cudaError_t launch_coopb(int NT, const OnchipParams& P, int sm_count, cudaStream_t st) {
  const bool sig = P.sigma != 0.0;
  switch (NT) {
#define MPCB_CB(N_) case N_: return sig ? launch_coopb_t<N_, true>(P, sm_count, st) : launch_coopb_t<N_, false>(P, sm_count, st);
    MPCB_CB(8) MPCB_CB(16) MPCB_CB(24) MPCB_CB(32) MPCB_CB(40) MPCB_CB(48) MPCB_CB(56) MPCB_CB(64) MPCB_CB(72) MPCB_CB(80) MPCB_CB(88) MPCB_CB(96) MPCB_CB(104) MPCB_CB(112) MPCB_CB(120)
#undef MPCB_CB
    default: return cudaErrorInvalidValue;
  }
}

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
