// C ABI of libmpcb200.so (see include/mpcb200.h).  Host plumbing only: design upload, buffer management, kernel
// dispatch, CUDA-event timing.  There is deliberately NO CPU fallback: without a usable sm_100 device every entry
// point that computes fails with MPCB_ERR_NO_DEVICE / MPCB_ERR_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mpcb200.h"
#include "api_common.hpp"
#include "admm_onchip.cuh"
#include "admm_riccati.cuh"
#include "admm_stream.cuh"
#include "closed_loop.cuh"
#include "host_design.hpp"
#include "multi_device.hpp"
#include "recover.cuh"

namespace mpcb {
// admm_smem.cu (own translation unit: 14 kernel instantiations compile in parallel with this file)
size_t smemk_bytes_host(int NT, int np, bool sig);
cudaError_t launch_smemk(int NT, const OnchipParams& P, int sm_count, int* attr_set, cudaStream_t st);
size_t smemg_bytes_host(int NT, int np, bool sig);
size_t coop_bytes_host(int NT, int np, bool sig);
cudaError_t launch_coop(int NT, const OnchipParams& P, int sm_count, cudaStream_t st);       // CTA-cooperative straggler kernel (admm_coop.cuh), NT = 24 .. 120
size_t coopb_bytes_host(int NT, int np, bool sig);
cudaError_t launch_coopb(int NT, const OnchipParams& P, int sm_count, cudaStream_t st);      // its box-only counterpart for small batches, NT = 8 .. 120
cudaError_t launch_smemg(int NT, const OnchipParams& P, int sm_count, cudaStream_t st);      // general rows, 64 < nt <= 120 (admm_smemg.cuh)
}  // namespace mpcb

namespace {
thread_local std::string g_err;
}  // namespace

namespace mpcb {
int api_fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
}  // namespace mpcb

namespace {
using mpcb::DevBuf;
using mpcb::PinBuf;
using mpcb::is_pinned_or_device;
using mpcb::upload;
int fail(int code, const std::string& msg) { return mpcb::api_fail(code, msg); }
}  // namespace

struct mpcb_handle {
  mpcb::Design D;
  mpcb::Design D2;      // second rung of the rho ladder (only T and rho_vec are used)
  mpcb_settings st;
  mpcb_info info;
  mpcb_timing timing;
  int NT = 0;  // padded operator size of the on-chip kernel
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;      // device-to-host leg of the pipelined host entry
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t chunk_ev[16] = {};
  // per-system constants
  DevBuf<double> Tfrag, Cfrag, Lt, Lv, lo, hi, rho, rinv, A, B, Q, R, S, P;
  DevBuf<double> ric_stage, ric_Lq;         // stage-wise (Riccati) kernel: per-stage factors, Lq row-major
  DevBuf<double> rW, rQ, rD, rX, rXT, rYO;  // its tile workspaces [tiles][nz][32]
  size_t smem_optin = 0;
  cudaEvent_t busy_ev = nullptr;            // end of the last enqueued call: the next call's stream waits on it (one call in flight per handle)
  bool busy_recorded = false;
  // multi-device handle (settings.n_devices > 1): this handle is the shard owner of device_ids[0]; `peers` own the other devices
  std::vector<mpcb_handle*> peers;
  cudaEvent_t fan_ev = nullptr, done_ev = nullptr;   // device entry: inputs ready on the caller's stream / this peer's shard is back in the caller's arrays
  DevBuf<double> Tfrag2, rho2, rinv2;       // second rung of the rho ladder (settings.ladder_iter): operator and step sizes
  DevBuf<int32_t> remap;                    // problems the first pass left unsolved
  bool ladder = false;
  DevBuf<unsigned long long> counter;
  mpcb::StreamConsts sc;  // streamed-kernel constants
  mpcb::StreamConsts sc2; // ... of the second rung of the rho ladder
  PinBuf<unsigned long long> ladder_count;
  // batch workspaces (device)
  DevBuf<double> x0, xref, uref, warm_v, warm_y, v, y, pres, dres, u, e_u, x, e_x, u0, obj;
  DevBuf<int32_t> status, iters;
  mpcb::StreamWork sw;
  // pinned staging for pageable user memory
  PinBuf<double> stage_in, stage_out;
  PinBuf<int32_t> stage_int;
  PinBuf<double> small_io;     // small batches: inputs and outputs live in one page-locked block ...
  DevBuf<double> small_dev;    // ... optionally mirrored in device memory (MPCB_SMALL_MIRROR: one copy in, one copy out)
  int onchip_blocks_per_sm = 0;
  size_t recover_smem_set = 0, recover_wide_smem_set = 0;
};

namespace {

using mpcb::OnchipParams;

// rho ladder: the indices of the problems that hit the first pass's iteration cap, in any order; count[2] of the work counter block
__global__ void collect_unsolved_kernel(const int32_t* status, long long n, int32_t* remap, unsigned long long* count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && status[i] == MPCB_STATUS_MAX_ITER) remap[atomicAdd(count, 1ULL)] = (int32_t)i;
}

template <int NT, bool HAS_G, bool SIG, int MINB>
cudaError_t launch_onchip_t(const OnchipParams& P, int sm_count, int* blocks_per_sm_cache, cudaStream_t st) {
  auto kern = mpcb::admm_onchip_kernel<NT, HAS_G, SIG, MINB>;
  const size_t smem = mpcb::onchip_smem_bytes(NT, P.np, HAS_G);
  if (*blocks_per_sm_cache == 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);      // the device maximum (see admm_smem.cu)
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, mpcb::ONCHIP_THREADS, smem);
    if (e != cudaSuccess) return e;
    if (const char* cap = std::getenv("MPCB_CTAS_PER_SM")) { const int c = std::atoi(cap); if (c > 0) occ = std::min(occ, c); }   // experiment knob (tools/ab_onchip.sh)
    *blocks_per_sm_cache = std::max(occ, 1);
  }
  const long long warps_needed = (P.batch + 7) / 8;
  const long long blocks_needed = (warps_needed + mpcb::ONCHIP_WARPS - 1) / mpcb::ONCHIP_WARPS;
  // A batch that gives each slot slightly more than one problem is the worst case of the dynamic queue: a few slots get a second
  // problem and everyone waits for them.  One CTA per SM fewer makes that ~2 problems for every slot.  Measured on B200 (16 384 QT
  // problems, 3 -> 2 CTAs/SM): H = 20 0.185 -> 0.164 ms, H = 30 0.438 -> 0.397 ms; no difference at 8 192 or 32 768 problems.
  int occ = *blocks_per_sm_cache;
  const double per_slot = (double)P.batch / ((double)sm_count * occ * mpcb::ONCHIP_WARPS * 8);
  if (occ >= 2 && per_slot > 1.0 && per_slot < 1.6) occ -= 1;
  const long long grid = std::min<long long>(blocks_needed, (long long)sm_count * occ);
  kern<<<(unsigned)std::max<long long>(grid, 1), mpcb::ONCHIP_THREADS, smem, st>>>(P);
  return cudaGetLastError();
}

// MINB (CTAs of 128 threads per SM) is the register budget of each variant (checked with -Xptxas -v: no spills for
// the box-only kernels; the general-row kernels above NT = 40 spill a few hundred bytes).
template <bool HAS_G, bool SIG>
cudaError_t launch_onchip_gs(int NT, const OnchipParams& P, int sm_count, int* cache, cudaStream_t st) {
  switch (NT) {
    case 8: return launch_onchip_t<8, HAS_G, SIG, 4>(P, sm_count, cache, st);
    case 16: return launch_onchip_t<16, HAS_G, SIG, 4>(P, sm_count, cache, st);
    case 24: return launch_onchip_t<24, HAS_G, SIG, HAS_G ? 3 : 4>(P, sm_count, cache, st);
    case 32: return launch_onchip_t<32, HAS_G, SIG, HAS_G ? 2 : 4>(P, sm_count, cache, st);
    case 40: return launch_onchip_t<40, HAS_G, SIG, HAS_G ? 2 : 3>(P, sm_count, cache, st);
    case 48: return launch_onchip_t<48, HAS_G, SIG, HAS_G ? 2 : 3>(P, sm_count, cache, st);      // box-only: fits 168 registers (3 CTAs/SM) without spills, sigma = 0 or not
    case 56: return launch_onchip_t<56, HAS_G, SIG, (HAS_G || SIG) ? 2 : 3>(P, sm_count, cache, st);      // sigma > 0 would spill at 168 registers
    case 64: return launch_onchip_t<64, HAS_G, SIG, (HAS_G || SIG) ? 2 : 3>(P, sm_count, cache, st);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_onchip(int NT, bool has_g, const OnchipParams& P, int sm_count, int* cache, cudaStream_t st) {
  if (has_g) return launch_onchip_gs<true, true>(NT, P, sm_count, cache, st);
  return P.sigma != 0.0 ? launch_onchip_gs<false, true>(NT, P, sm_count, cache, st) : launch_onchip_gs<false, false>(NT, P, sm_count, cache, st);
}

// fragment order of a symmetric NT x NT operator for the on-chip kernel (see admm_onchip.cuh)
std::vector<double> to_fragments(const mpcb::Mat& M, int nt, int NT) {
  const int KS = NT / 4, NTL = NT / 8;
  std::vector<double> f((size_t)NT * NT, 0.0);
  for (int s = 0; s < KS; s++)
    for (int tn = 0; tn < NTL; tn++)
      for (int lane = 0; lane < 32; lane++) {
        const int g = lane >> 2, l4 = lane & 3;
        const int row = 8 * (s >> 1) + 2 * l4 + (s & 1), col = 8 * tn + g;
        f[((size_t)s * NTL + tn) * 32 + lane] = (row < nt && col < nt) ? M(row, col) : 0.0;
      }
  return f;
}

int upload_design(mpcb_handle* h) {
  const mpcb::Design& D = h->D;
  CUDA_TRY(upload(h->A, D.A.a.data(), D.A.a.size()));
  CUDA_TRY(upload(h->B, D.B.a.data(), D.B.a.size()));
  CUDA_TRY(upload(h->Q, D.Q.a.data(), D.Q.a.size()));
  CUDA_TRY(upload(h->R, D.R.a.data(), D.R.a.size()));
  CUDA_TRY(upload(h->S, D.S.a.data(), D.S.a.size()));
  CUDA_TRY(upload(h->P, D.P.a.data(), D.P.a.size()));
  CUDA_TRY(h->counter.ensure(3));
  CUDA_TRY(cudaMemset(h->counter.p, 0, 3 * sizeof(unsigned long long)));   // once: the on-chip kernel re-arms it itself
  if (h->info.kernel == MPCB_KERNEL_RICCATI) {
    std::vector<double> lq((size_t)D.nz * D.np);
    for (int i = 0; i < D.nz; i++)
      for (int j = 0; j < D.np; j++) lq[(size_t)i * D.np + j] = D.Lq(i, j);
    CUDA_TRY(upload(h->ric_stage, D.ric_stage.data(), D.ric_stage.size()));
    CUDA_TRY(upload(h->ric_Lq, lq.data(), lq.size()));
    if (!D.Lv.a.empty()) {
      for (int i = 0; i < D.nz; i++)
        for (int j = 0; j < D.np; j++) lq[(size_t)i * D.np + j] = D.Lv(i, j);
      CUDA_TRY(upload(h->Lv, lq.data(), lq.size()));
    }
  } else if (h->info.kernel == MPCB_KERNEL_ONCHIP || h->info.kernel == MPCB_KERNEL_ONCHIP_SMEM) {
    const int NT = h->NT, nt = D.nt, np = D.np;
    std::vector<double> tf = to_fragments(D.T, nt, NT), cf = to_fragments(D.C, nt, NT);
    std::vector<double> Lt((size_t)np * NT, 0.0), lo(NT, 0.0), hi(NT, 0.0), rho(NT, 1.0), rinv(NT, 1.0);
    for (int j = 0; j < np; j++) {
      for (int i = 0; i < D.nz; i++) Lt[(size_t)j * NT + i] = D.Lq(i, j);
      for (int i = 0; i < D.mg; i++) Lt[(size_t)j * NT + D.nz + i] = D.Lb(i, j);
    }
    for (int i = 0; i < nt; i++) { lo[i] = D.lo[i]; hi[i] = D.hi[i]; rho[i] = D.rho_vec[i]; rinv[i] = 1.0 / D.rho_vec[i]; }
    CUDA_TRY(upload(h->Tfrag, tf.data(), tf.size()));
    CUDA_TRY(upload(h->Cfrag, cf.data(), cf.size()));
    CUDA_TRY(upload(h->Lt, Lt.data(), Lt.size()));
    if (!D.Lv.a.empty()) {      // cold-start map in the same [np][NT] layout (general rows stay zero)
      std::fill(Lt.begin(), Lt.end(), 0.0);
      for (int j = 0; j < np; j++)
        for (int i = 0; i < D.nz; i++) Lt[(size_t)j * NT + i] = D.Lv(i, j);
      CUDA_TRY(upload(h->Lv, Lt.data(), Lt.size()));
    }
    CUDA_TRY(upload(h->lo, lo.data(), NT));
    CUDA_TRY(upload(h->hi, hi.data(), NT));
    CUDA_TRY(upload(h->rho, rho.data(), NT));
    CUDA_TRY(upload(h->rinv, rinv.data(), NT));
    if (h->ladder) {
      std::vector<double> tf2 = to_fragments(h->D2.T, nt, NT);
      for (int i = 0; i < nt; i++) { rho[i] = h->D2.rho_vec[i]; rinv[i] = 1.0 / h->D2.rho_vec[i]; }
      CUDA_TRY(upload(h->Tfrag2, tf2.data(), tf2.size()));
      CUDA_TRY(upload(h->rho2, rho.data(), NT));
      CUDA_TRY(upload(h->rinv2, rinv.data(), NT));
    }
  } else {
    std::string err;
    cudaError_t e = mpcb::stream_upload(D, h->sc, err);
    if (e != cudaSuccess) return fail(MPCB_ERR_CUDA, "stream_upload: " + err + cudaGetErrorString(e));
    if (h->ladder) {
      e = mpcb::stream_upload(h->D2, h->sc2, err);
      if (e != cudaSuccess) return fail(MPCB_ERR_CUDA, "stream_upload (second rung): " + err + cudaGetErrorString(e));
    }
  }
  return MPCB_OK;
}

// Enqueue solve + recover on `st` with device-resident io.  Internal scratch is used for anything the caller skipped.
// `v_keep` (optional): device buffer [batch][nz] that receives the absolute inputs instead of the handle's scratch.
int enqueue_device(mpcb_handle* h, const mpcb_batch_io& io, cudaStream_t st, cudaEvent_t ev_mid, double* v_keep = nullptr) {
  const mpcb::Design& D = h->D;
  const long long Bn = io.batch;
  if (Bn <= 0) return fail(MPCB_ERR_INVALID, "batch must be positive");
  if (!io.x0 || !io.xref || !io.uref) return fail(MPCB_ERR_INVALID, "x0, xref, uref are required");
  if ((io.warm_u == nullptr) != (io.warm_y == nullptr)) return fail(MPCB_ERR_INVALID, "warm_u and warm_y must be given together");
  double* v_buf = v_keep;
  if (!v_buf) { CUDA_TRY(h->v.ensure((size_t)Bn * D.nz)); v_buf = h->v.p; }
  int32_t* d_status = io.status; int32_t* d_iters = io.iters; double* d_pres = io.prim_res; double* d_dres = io.dual_res;
  if (!d_status) { CUDA_TRY(h->status.ensure(Bn)); d_status = h->status.p; }
  if (!d_iters) { CUDA_TRY(h->iters.ensure(Bn)); d_iters = h->iters.p; }
  if (!d_pres) { CUDA_TRY(h->pres.ensure(Bn)); d_pres = h->pres.p; }
  if (!d_dres) { CUDA_TRY(h->dres.ensure(Bn)); d_dres = h->dres.p; }
  int launches = 0;
  // One call in flight per handle (work queue counter, scratch buffers and workspaces are per handle): a call enqueued on another
  // stream while the previous one still runs is ordered behind it instead of racing with it.
  if (h->busy_recorded) CUDA_TRY(cudaStreamWaitEvent(st, h->busy_ev, 0));
  if (h->info.kernel == MPCB_KERNEL_RICCATI) {
    const size_t tile_doubles = (size_t)((Bn + 31) / 32) * D.nz * 32;
    const bool sig = h->st.sigma != 0.0;
    CUDA_TRY(h->rW.ensure(tile_doubles)); CUDA_TRY(h->rQ.ensure(tile_doubles)); CUDA_TRY(h->rD.ensure(tile_doubles)); CUDA_TRY(h->rXT.ensure(tile_doubles));
    if (sig) CUDA_TRY(h->rX.ensure(tile_doubles));
    if (io.y) CUDA_TRY(h->rYO.ensure(tile_doubles));
    mpcb::RiccatiParams P{};
    P.stage = h->ric_stage.p; P.Lq = h->ric_Lq.p; P.Lv = h->st.cold_init ? h->Lv.p : nullptr;
    for (int m = 0; m < D.nx; m++)
      for (int i = 0; i < D.nu; i++) { P.Bm[m * D.nu + i] = D.B(m, i); P.Bt[i * D.nx + m] = D.B(m, i); }
    for (int i = 0; i < D.nu; i++) { P.lo[i] = D.lo[i]; P.hi[i] = D.hi[i]; }
    P.H = D.H; P.nz = D.nz; P.np = D.np; P.ch = 0;
    P.rho = D.rho; P.sigma = h->st.sigma; P.alpha = h->st.alpha; P.eps_abs = h->st.eps_abs; P.eps_rel = h->st.eps_rel;
    P.max_iter = h->st.max_iter; P.check_every = h->st.check_every;
    P.batch = Bn; P.x0 = io.x0; P.xref = io.xref; P.uref = io.uref; P.xref_bc = io.xref_broadcast; P.uref_bc = io.uref_broadcast;
    P.warm_v = io.warm_u; P.warm_y = io.warm_y; P.v_out = v_buf; P.y_out = io.y;
    P.status = d_status; P.iters = d_iters; P.pres = d_pres; P.dres = d_dres;
    P.W = h->rW.p; P.Qb = h->rQ.p; P.D = h->rD.p; P.X = sig ? h->rX.p : nullptr; P.XT = h->rXT.p; P.YO = io.y ? h->rYO.p : nullptr;
    P.counter = h->counter.p;
    cudaError_t e = mpcb::launch_riccati(D.nx, D.nu, sig, P, h->info.sm_count, h->smem_optin, st);
    if (e != cudaSuccess) return fail(MPCB_ERR_CUDA, std::string("admm_riccati launch: ") + cudaGetErrorString(e));
    launches += 1;
  } else if (h->info.kernel == MPCB_KERNEL_ONCHIP || h->info.kernel == MPCB_KERNEL_ONCHIP_SMEM) {
    OnchipParams P{};
    P.Tfrag = h->Tfrag.p; P.Cfrag = h->Cfrag.p; P.Lt = h->Lt.p; P.Lv = h->st.cold_init ? h->Lv.p : nullptr; P.lo = h->lo.p; P.hi = h->hi.p; P.rho = h->rho.p; P.rinv = h->rinv.p;
    P.nz = D.nz; P.nt = D.nt; P.np = D.np; P.nx = D.nx; P.nu = D.nu; P.nball = D.nball;
    P.rho_box = D.rho; P.sigma = h->st.sigma; P.alpha = h->st.alpha; P.eps_abs = h->st.eps_abs; P.eps_rel = h->st.eps_rel;
    P.eps_pinf = h->st.eps_prim_inf; P.max_iter = h->ladder ? h->st.ladder_iter : h->st.max_iter; P.check_every = h->st.check_every;
    P.batch = Bn; P.x0 = io.x0; P.xref = io.xref; P.uref = io.uref; P.xref_bc = io.xref_broadcast; P.uref_bc = io.uref_broadcast;
    P.warm_v = io.warm_u; P.warm_y = io.warm_y; P.v_out = v_buf; P.y_out = io.y;
    if (h->ladder && !P.y_out) { CUDA_TRY(h->y.ensure((size_t)Bn * D.nt)); P.y_out = h->y.p; }     // the second rung starts from the first pass's iterate
    P.status = d_status; P.iters = d_iters; P.pres = d_pres; P.dres = d_dres; P.counter = h->counter.p;
    auto launch_slots = [&](const OnchipParams& Q) {      // register-resident (nt <= 64) or shared-memory resident (box-only / general rows)
      if (h->info.kernel == MPCB_KERNEL_ONCHIP) return launch_onchip(h->NT, D.mg > 0, Q, h->info.sm_count, &h->onchip_blocks_per_sm, st);
      if (D.mg > 0) return mpcb::launch_smemg(h->NT, Q, h->info.sm_count, st);
      return mpcb::launch_smemk(h->NT, Q, h->info.sm_count, &h->onchip_blocks_per_sm, st);
    };
    // A SMALL batch on a controller with general rows (the closed-loop, few-plants-at-a-time use of update_initialization! / calculate!) is one
    // warp's latency on the slot kernels; the CTA-cooperative kernel gives every group of eight problems a whole CTA (bit-identical results).
    const bool coop_ok = D.mg > 0 && D.nball == 0 && h->NT >= 24 && h->NT <= 120 && mpcb::coop_bytes_host(h->NT, D.np, h->st.sigma != 0.0) <= h->smem_optin;
    static const bool no_small_coop = std::getenv("MPCB_NO_SMALL_COOP") != nullptr;      // A/B switch for measurements
    const bool coopb_ok = D.mg == 0 && h->NT <= 120 && mpcb::coopb_bytes_host(h->NT, D.np, h->st.sigma != 0.0) <= h->smem_optin;      // box-only counterpart
    // How small is small (profiles/r02/small_threshold_probe.jsonl, box-only QT, device-resident, solve + recover): against the register-resident slot kernel
    // (NT <= 64) the cooperative kernel wins up to 8 problems per SM (H = 20: 74 vs 93 us at 1 184 problems, 128 vs 105 us at 2 368); against the
    // shared-memory resident ones (NT >= 72: one CTA of 4 .. 8 warps per SM, each warp alone on its scheduler) up to 32 per SM (H = 50: 183 vs 638 us at
    // 1 184 problems, 333 vs 669 at 2 368, 585 vs 668 at 4 736, 1 081 vs 844 at 9 472).
    static const long long small_env = [] { const char* e = std::getenv("MPCB_SMALL_PER_SM"); return e ? std::atoll(e) : 0LL; }();      // A/B knob
    const long long small_per_sm = small_env > 0 ? small_env : (h->NT >= 72 ? 32LL : 8LL);
    const bool small_batch = (coop_ok || coopb_ok) && !no_small_coop && Bn <= small_per_sm * h->info.sm_count;
    cudaError_t e;
    if (small_batch && D.mg > 0) { OnchipParams Pc = P; Pc.tickets_max = -1; e = mpcb::launch_coop(h->NT, Pc, h->info.sm_count, st); }
    else if (small_batch) e = mpcb::launch_coopb(h->NT, P, h->info.sm_count, st);
    else e = launch_slots(P);
    if (e != cudaSuccess) return fail(MPCB_ERR_CUDA, std::string("admm_onchip launch: ") + cudaGetErrorString(e));
    launches += 1;
    if (h->ladder) {
      // second rung: the problems that ran into the first pass's cap, warm-started at the first pass's (x, y) -- duals do not depend
      // on rho -- with the stiffer state-box step sizes (oracle experiment: warm 335 mean / 915 max iterations, cold 646 / 1260).
      // No host round trip: the kernel reads the number of tickets from the counter block; with none, its CTAs exit at once.
      CUDA_TRY(h->remap.ensure(Bn));
      CUDA_TRY(cudaMemsetAsync(h->counter.p + 2, 0, sizeof(unsigned long long), st));
      collect_unsolved_kernel<<<(unsigned)((Bn + 255) / 256), 256, 0, st>>>(d_status, Bn, h->remap.p, h->counter.p + 2);
      CUDA_TRY(cudaGetLastError());
      OnchipParams P2 = P;
      P2.Tfrag = h->Tfrag2.p; P2.rho = h->rho2.p; P2.rinv = h->rinv2.p;
      P2.max_iter = h->st.max_iter - h->st.ladder_iter; P2.iters_add = P.max_iter;
      P2.warm_v = P.v_out; P2.warm_y = P.y_out;
      P2.remap = h->remap.p; P2.batch_dev = h->counter.p + 2;
      // A SMALL second rung (the usual case: a few problems per 10^3) is latency, not throughput: the CTA-cooperative kernel gives every group of
      // eight stragglers a whole CTA.  The count lives on the device, so both kernels are enqueued and each looks at it: up to four groups per SM
      // go to the cooperative kernel, anything larger to the slot kernel.
      if (coop_ok) {
        OnchipParams Pc = P2;
        Pc.tickets_max = 32LL * h->info.sm_count;
        e = mpcb::launch_coop(h->NT, Pc, h->info.sm_count, st);
        if (e != cudaSuccess) return fail(MPCB_ERR_CUDA, std::string("admm_coop (second rung) launch: ") + cudaGetErrorString(e));
        P2.tickets_skip_le = Pc.tickets_max;
        launches += 1;
      }
      e = launch_slots(P2);
      if (e != cudaSuccess) return fail(MPCB_ERR_CUDA, std::string("admm_onchip (second rung) launch: ") + cudaGetErrorString(e));
      launches += 2;
    }
  } else {
    std::string err;
    mpcb::StreamBatch sb;
    sb.batch = Bn; sb.x0 = io.x0; sb.xref = io.xref; sb.uref = io.uref; sb.xref_bc = io.xref_broadcast; sb.uref_bc = io.uref_broadcast;
    sb.warm_v = io.warm_u; sb.warm_y = io.warm_y; sb.v_out = v_buf; sb.y_out = io.y;
    sb.status = d_status; sb.iters = d_iters; sb.pres = d_pres; sb.dres = d_dres;
    int nl = 0;
    mpcb_settings st1 = h->st;
    if (h->ladder) {        // first rung: capped; its (x, y) are the second rung's warm start
      st1.max_iter = h->st.ladder_iter;
      if (!sb.y_out) { CUDA_TRY(h->y.ensure((size_t)Bn * D.nt)); sb.y_out = h->y.p; }
    }
    cudaError_t e = mpcb::stream_solve(h->D, st1, h->sc, h->sw, sb, h->info.sm_count, st, &nl, err);
    if (e != cudaSuccess) return fail(MPCB_ERR_CUDA, "admm_stream: " + err + " " + cudaGetErrorString(e));
    launches += nl;
    if (h->ladder) {
      // second rung (as on the on-chip kernel): the problems that ran into the cap continue from their iterate with the stiffer state-box
      // step sizes of the second cached operator.  The streamed path is host-driven anyway, so the count is simply read back.
      CUDA_TRY(h->remap.ensure(Bn)); CUDA_TRY(h->ladder_count.ensure(1));
      CUDA_TRY(cudaMemsetAsync(h->counter.p + 2, 0, sizeof(unsigned long long), st));
      collect_unsolved_kernel<<<(unsigned)((Bn + 255) / 256), 256, 0, st>>>(d_status, Bn, h->remap.p, h->counter.p + 2);
      CUDA_TRY(cudaGetLastError());
      CUDA_TRY(cudaMemcpyAsync(h->ladder_count.p, h->counter.p + 2, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaStreamSynchronize(st));
      launches += 1;
      const long long n2 = (long long)h->ladder_count.p[0];
      if (n2 > 0) {
        mpcb::StreamBatch s2 = sb;
        s2.batch = n2; s2.remap = h->remap.p; s2.iters_add = h->st.ladder_iter;
        s2.warm_v = v_buf; s2.warm_y = sb.y_out;           // in place: a row reads its problem's warm start in the init kernel, before anything is written back
        mpcb_settings st2 = h->st;
        st2.max_iter = h->st.max_iter - h->st.ladder_iter;
        int nl2 = 0;
        e = mpcb::stream_solve(h->D2, st2, h->sc2, h->sw, s2, h->info.sm_count, st, &nl2, err);
        if (e != cudaSuccess) return fail(MPCB_ERR_CUDA, "admm_stream (second rung): " + err + " " + cudaGetErrorString(e));
        launches += nl2;
      }
    }
  }
  if (ev_mid) CUDA_TRY(cudaEventRecord(ev_mid, st));
  if (io.u || io.e_u || io.x || io.e_x || io.u0 || io.objective) {
    mpcb::RecoverParams R;
    R.A = h->A.p; R.B = h->B.p; R.Q = h->Q.p; R.R = h->R.p; R.S = h->S.p; R.Pt = h->P.p;
    R.nx = D.nx; R.nu = D.nu; R.H = D.H; R.use_R = D.use_R; R.use_S = D.use_S; R.batch = Bn;
    R.x0 = io.x0; R.xref = io.xref; R.uref = io.uref; R.xref_bc = io.xref_broadcast; R.uref_bc = io.uref_broadcast;
    R.v = v_buf; R.u = io.u; R.e_u = io.e_u; R.x = io.x; R.e_x = io.e_x; R.u0 = io.u0; R.objective = io.objective;
    if (mpcb::launch_recover_small(R, st)) {
    } else if (mpcb::recover_wide_applies(D.nx, D.nu)) {
      bool qdiag = true;
      for (int j = 0; j < D.nx && qdiag; j++)
        for (int i = 0; i < D.nx; i++) if (i != j && D.Q(i, j) != 0.0) { qdiag = false; break; }
      const void* kern = mpcb::recover_wide_variant(D.nx, qdiag);
      const size_t smem = mpcb::recover_wide_smem_bytes(D.nx, D.nu);
      if (smem > 48 * 1024 && smem > h->recover_wide_smem_set) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        h->recover_wide_smem_set = smem;
      }
      const int tp = ((D.nx + 31) / 32) * 32, gpc = mpcb::RECOVER_WIDE_THREADS / tp;
      int per_sm = 1;
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, mpcb::RECOVER_WIDE_THREADS, smem));
      const long long need = (Bn + gpc - 1) / gpc;
      const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(need, (long long)h->info.sm_count * std::max(per_sm, 1)));
      mpcb::RecoverParams Rc = R;
      void* args[] = {(void*)&Rc};
      CUDA_TRY(cudaLaunchKernel(kern, dim3(grid), dim3(mpcb::RECOVER_WIDE_THREADS), args, smem, st));
    } else {
      const int rt = mpcb::recover_threads_for(D.nx, D.nu);
      if (rt == 0) return fail(MPCB_ERR_INVALID, "result recovery: system too wide for the generic kernel (nx, nu)");
      const size_t smem = mpcb::recover_smem_bytes(D.nx, D.nu, rt);
      if (smem > 48 * 1024 && smem > h->recover_smem_set) {
        CUDA_TRY(cudaFuncSetAttribute(mpcb::recover_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        h->recover_smem_set = smem;
      }
      const unsigned grid = (unsigned)((Bn + rt - 1) / rt);
      mpcb::recover_kernel<<<grid, rt, smem, st>>>(R);
    }
    CUDA_TRY(cudaGetLastError());
    launches += 1;
  }
  CUDA_TRY(cudaEventRecord(h->busy_ev, st));
  h->busy_recorded = true;
  h->timing.kernel_launches = launches;
  h->timing.batch = Bn;
  return MPCB_OK;
}

mpcb::ShardDims dims_of(const mpcb_handle* h) {
  const mpcb::Design& D = h->D;
  return mpcb::ShardDims{(size_t)D.nx, (size_t)D.nu, (size_t)D.H, (size_t)D.nz, (size_t)D.nt};
}

// Device entry of a multi-device handle: the caller's buffers live on device_ids[0].  Shard r > 0 travels to its device as peer copies
// (NVLink) on that device's stream, is solved there, and its results are peer-copied straight back into the caller's arrays; the caller's
// stream waits on every peer's completion event.  Everything is enqueued: the call stays asynchronous with respect to the host (for the
// on-chip / stage-wise kernels).
int enqueue_device_multi(mpcb_handle* h, const mpcb_batch_io& io, cudaStream_t st) {
  const int ndev = 1 + (int)h->peers.size();
  const mpcb::ShardDims d = dims_of(h);
  const long long Bn = io.batch;
  const int root = h->st.device;
  CUDA_TRY(cudaEventRecord(h->fan_ev, st));
  int launches = 0;
  for (int r = 1; r < ndev; r++) {
    mpcb_handle* p = h->peers[(size_t)r - 1];
    long long lo, hi;
    mpcb::shard_range(Bn, r, ndev, &lo, &hi);
    const long long n = hi - lo;
    if (n <= 0) continue;
    const int dev = p->st.device;
    CUDA_TRY(cudaSetDevice(dev));
    cudaStream_t ps = p->stream;
    CUDA_TRY(cudaStreamWaitEvent(ps, h->fan_ev, 0));
    const mpcb_batch_io sio = mpcb::shard_io(io, lo, n, d);      // this shard's slices of the caller's (root-device) arrays
    mpcb_batch_io pio = sio;                                       // the same shard in this peer's own buffers
    struct In { const double* src; DevBuf<double>* dst; size_t cnt; const double** slot; };
    In ins[5] = {{sio.x0, &p->x0, d.nx * (size_t)n, &pio.x0}, {sio.xref, &p->xref, io.xref_broadcast ? d.nx : d.nx * (size_t)n, &pio.xref},
                 {sio.uref, &p->uref, io.uref_broadcast ? d.nu : d.nu * (size_t)n, &pio.uref}, {sio.warm_u, &p->warm_v, d.nz * (size_t)n, &pio.warm_u},
                 {sio.warm_y, &p->warm_y, d.ny * (size_t)n, &pio.warm_y}};
    for (auto& in : ins) {
      if (!in.src) continue;
      CUDA_TRY(in.dst->ensure(in.cnt));
      CUDA_TRY(cudaMemcpyPeerAsync(in.dst->p, dev, in.src, root, in.cnt * sizeof(double), ps));
      *in.slot = in.dst->p;
    }
    struct Out { double* dst; DevBuf<double>* buf; size_t cnt; double** slot; };
    Out outs[9] = {{sio.u, &p->u, d.nz * (size_t)n, &pio.u}, {sio.e_u, &p->e_u, d.nz * (size_t)n, &pio.e_u}, {sio.x, &p->x, d.nx * (d.H + 1) * (size_t)n, &pio.x},
                   {sio.e_x, &p->e_x, d.nx * (d.H + 1) * (size_t)n, &pio.e_x}, {sio.u0, &p->u0, d.nu * (size_t)n, &pio.u0}, {sio.prim_res, &p->pres, (size_t)n, &pio.prim_res},
                   {sio.dual_res, &p->dres, (size_t)n, &pio.dual_res}, {sio.objective, &p->obj, (size_t)n, &pio.objective}, {sio.y, &p->y, d.ny * (size_t)n, &pio.y}};
    for (auto& o : outs) {
      if (!o.dst) continue;
      CUDA_TRY(o.buf->ensure(o.cnt));
      *o.slot = o.buf->p;
    }
    CUDA_TRY(p->status.ensure((size_t)n)); CUDA_TRY(p->iters.ensure((size_t)n));
    pio.status = p->status.p; pio.iters = p->iters.p;
    int rc = enqueue_device(p, pio, ps, nullptr);
    if (rc != MPCB_OK) return rc;
    launches += p->timing.kernel_launches;
    for (auto& o : outs)
      if (o.dst) CUDA_TRY(cudaMemcpyPeerAsync(o.dst, root, o.buf->p, dev, o.cnt * sizeof(double), ps));
    if (sio.status) CUDA_TRY(cudaMemcpyPeerAsync(sio.status, root, p->status.p, dev, (size_t)n * sizeof(int32_t), ps));
    if (sio.iters) CUDA_TRY(cudaMemcpyPeerAsync(sio.iters, root, p->iters.p, dev, (size_t)n * sizeof(int32_t), ps));
    CUDA_TRY(cudaEventRecord(p->done_ev, ps));
  }
  CUDA_TRY(cudaSetDevice(root));
  long long lo0, hi0;
  mpcb::shard_range(Bn, 0, ndev, &lo0, &hi0);
  int rc = enqueue_device(h, mpcb::shard_io(io, lo0, hi0 - lo0, d), st, nullptr);
  if (rc != MPCB_OK) return rc;
  launches += h->timing.kernel_launches;
  for (mpcb_handle* p : h->peers) CUDA_TRY(cudaStreamWaitEvent(st, p->done_ev, 0));
  h->timing.kernel_launches = launches;
  h->timing.batch = Bn;
  return MPCB_OK;
}

int solve_linear_host(mpcb_handle* h, const mpcb_batch_io* hio);

}  // namespace

extern "C" {

int mpcb_version(void) { return MPCB_VERSION; }

int mpcb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char* mpcb_last_error(void) { return g_err.c_str(); }

void mpcb_default_settings(mpcb_settings* s) {
  if (!s) return;
  std::memset(s, 0, sizeof(*s));
  s->eps_abs = 1e-3; s->eps_rel = 1e-3; s->eps_prim_inf = 1e-4; s->rho = 0.0; s->rho_eq_scale = 1e3;
  s->sigma = 1e-6; s->alpha = 1.6; s->max_iter = 4000; s->check_every = 25; s->device = 0; s->kernel = MPCB_KERNEL_AUTO; s->cold_init = 0;
}

void* mpcb_alloc_pinned(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { g_err = "cudaMallocHost failed"; cudaGetLastError(); return nullptr; }
  return p;
}
void mpcb_free_pinned(void* p) { if (p) cudaFreeHost(p); }

int mpcb_dare(int32_t nx, int32_t nu, const double* A, const double* B, const double* Q, const double* R, double* P_out) {
  if (nx <= 0 || nu <= 0 || !A || !B || !Q || !R || !P_out) return fail(MPCB_ERR_INVALID, "mpcb_dare: bad arguments");
  mpcb::Mat P;
  std::string err;
  if (!mpcb::dare_sda(mpcb::Mat::from(A, nx, nx), mpcb::Mat::from(B, nx, nu), mpcb::Mat::from(Q, nx, nx), mpcb::Mat::from(R, nu, nu), P, err))
    return fail(MPCB_ERR_NUMERIC, err);
  std::memcpy(P_out, P.a.data(), sizeof(double) * nx * nx);
  return MPCB_OK;
}

int mpcb_create_linear(const mpcb_linear_desc* desc, const mpcb_settings* settings, mpcb_handle** out) {
  if (!desc || !out) return fail(MPCB_ERR_INVALID, "mpcb_create_linear: null argument");
  *out = nullptr;
  mpcb_settings st;
  if (settings) st = *settings; else mpcb_default_settings(&st);
  if (st.check_every <= 0 || st.max_iter <= 0 || !(st.alpha > 0 && st.alpha < 2) || !(st.sigma >= 0) || !(st.eps_abs >= 0) || !(st.eps_rel >= 0))
    return fail(MPCB_ERR_INVALID, "mpcb_create_linear: invalid settings");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(MPCB_ERR_NO_DEVICE, "no CUDA device visible: libmpcb200 has no CPU fallback");
  }
  std::vector<int> dev_ids;
  { int rcd = mpcb::parse_devices(st, ndev, dev_ids); if (rcd != MPCB_OK) return rcd; }
  if (!dev_ids.empty()) st.device = dev_ids[0];
  if (st.device < 0 || st.device >= ndev) return fail(MPCB_ERR_INVALID, "settings.device out of range");
  CUDA_TRY(cudaSetDevice(st.device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, st.device));
  if (prop.major < 10) return fail(MPCB_ERR_NO_DEVICE, std::string("device ") + prop.name + " is not sm_100: this library is built for B200 only");

  mpcb_handle* h = new mpcb_handle();
  std::string err;
  int rc = mpcb::build_design(*desc, st, h->D, err);
  if (rc != MPCB_OK) { delete h; return fail(rc, err); }
  h->st = st;
  std::memset(&h->info, 0, sizeof(h->info));
  std::memset(&h->timing, 0, sizeof(h->timing));
  const mpcb::Design& D = h->D;
  h->info.nx = D.nx; h->info.nu = D.nu; h->info.horizon = D.H; h->info.nz = D.nz; h->info.mg = D.mg; h->info.nt = D.nt;
  h->info.rho = D.rho; h->info.lambda_min = D.lmin; h->info.lambda_max = D.lmax;
  h->info.device = st.device; h->info.sm_count = prop.multiProcessorCount;
  int kernel = st.kernel;
  const int nt8 = ((D.nt + 7) / 8) * 8;
  // The general-row variant of the shared-memory kernel also takes nt8 = 56 .. 64 by default.  Measured (profiles/r02/smemg_ab_v2_*.jsonl, 65 536
  // problems, terminal equality, register-resident vs shared-memory resident): nt8 = 40 0.650 vs 0.737 ms, 48: 0.897 vs 0.951 (nt = 44) and 1.288 vs
  // 1.210 (nt = 48) -- a toss-up, and the register kernel has the shorter single-problem latency --, 56: 1.93 vs 1.58, 64: 2.08 vs 1.83 ms; state
  // box H = 10 (nt = 60): 36.1 vs 27.6 ms, with the ladder 7.8 vs 6.0 ms.  kernel = 3 reaches it from nt8 = 32.
  const bool gen_rows = D.mg > 0;
  const int smem_lo = gen_rows ? (st.kernel == MPCB_KERNEL_ONCHIP_SMEM ? 32 : 56) : 72;
  const bool smem_ok = nt8 >= smem_lo && nt8 <= 120 &&
                       (D.mg == 0 ? mpcb::smemk_bytes_host(nt8, D.np, st.sigma != 0.0) : mpcb::smemg_bytes_host(nt8, D.np, st.sigma != 0.0)) <= (size_t)prop.sharedMemPerBlockOptin;
  h->smem_optin = (size_t)prop.sharedMemPerBlockOptin;
  const bool ric_form = !D.ric_stage.empty() && mpcb::riccati_supported(D.nx, D.nu);
  bool ric_fits = false;
  if (ric_form) { int wpc = 0, ch = 0; size_t sm = 0; ric_fits = mpcb::riccati_plan(D.nx, D.nu, D.H, st.sigma != 0.0, 1 << 20, prop.multiProcessorCount, h->smem_optin, &wpc, &ch, &sm); }
  // Kernel choice (measured crossover, profiles/r02/ricsweep_config4_*.jsonl, quadruple tank, 16 384 problems): the register- and shared-memory
  // resident DMMA kernels win while the condensed operator fits on chip (nt <= 120: H = 50 1.17 ms vs 2.70 ms stage-wise); beyond that the
  // stage-wise kernel replaces the streamed GEMM wherever its form applies (H = 75: 4.4 vs 6.7 ms, H = 200: 14.9 vs 33.9 ms).
  if (kernel == MPCB_KERNEL_AUTO) {
    if (gen_rows && smem_ok) kernel = MPCB_KERNEL_ONCHIP_SMEM;
    else kernel = (D.nt <= 64) ? MPCB_KERNEL_ONCHIP : (smem_ok ? MPCB_KERNEL_ONCHIP_SMEM : ((ric_form && ric_fits) ? MPCB_KERNEL_RICCATI : MPCB_KERNEL_STREAMED));
  }
  if (kernel == MPCB_KERNEL_RICCATI) {
    if (!ric_form) { delete h; return fail(MPCB_ERR_INVALID, "stage-wise (Riccati) kernel: box-only problems without the S term, (nx, nu) in the compiled set"); }
    if (!ric_fits) { delete h; return fail(MPCB_ERR_INVALID, "stage-wise (Riccati) kernel: the stage matrices of this horizon do not fit shared memory"); }
  }
  if (kernel == MPCB_KERNEL_ONCHIP && D.nt > 64) { delete h; return fail(MPCB_ERR_INVALID, "on-chip kernel needs nz + mg <= 64"); }
  if (D.nball > 0 && kernel != MPCB_KERNEL_ONCHIP && kernel != MPCB_KERNEL_ONCHIP_SMEM) { delete h; return fail(MPCB_ERR_INVALID, "the contractive terminal set is implemented in the on-chip kernels only: needs nz + mg <= 120"); }
  if (kernel == MPCB_KERNEL_ONCHIP_SMEM && !smem_ok) { delete h; return fail(MPCB_ERR_INVALID, "shared-memory kernel needs 64 < nz <= 120 (box-only) or 32 <= nz + mg <= 120 (general rows) and an operator that fits 227 KB"); }
  if (kernel != MPCB_KERNEL_ONCHIP && kernel != MPCB_KERNEL_STREAMED && kernel != MPCB_KERNEL_ONCHIP_SMEM && kernel != MPCB_KERNEL_RICCATI) { delete h; return fail(MPCB_ERR_INVALID, "unknown kernel id"); }
  h->info.kernel = kernel;
  // rho ladder: only where it applies -- inequality general rows (state box; the ball rows are not boxes) on the on-chip kernel
  if (st.ladder_iter > 0 && (kernel == MPCB_KERNEL_ONCHIP || kernel == MPCB_KERNEL_ONCHIP_SMEM || kernel == MPCB_KERNEL_STREAMED) && D.mg > D.nball && desc->state_constraint) {
    if (st.ladder_iter >= st.max_iter) { delete h; return fail(MPCB_ERR_INVALID, "settings.ladder_iter must be below max_iter"); }
    h->st.ladder_iter = ((st.ladder_iter + st.check_every - 1) / st.check_every) * st.check_every;
    if (h->st.ladder_kappa <= 0) h->st.ladder_kappa = 10;
    rc = mpcb::build_design(*desc, st, h->D2, err, (double)h->st.ladder_kappa);
    if (rc != MPCB_OK) { delete h; return fail(rc, "second rung of the rho ladder: " + err); }
    h->ladder = true;
  }
  h->NT = (kernel == MPCB_KERNEL_STREAMED) ? mpcb::stream_padded(D.nt) : (kernel == MPCB_KERNEL_RICCATI ? D.nt : nt8);
  h->info.nt_pad = h->NT;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return fail(MPCB_ERR_CUDA, "cudaStreamCreate failed"); }
  if (cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { mpcb_destroy(h); return fail(MPCB_ERR_CUDA, "cudaStreamCreate failed"); }
  for (auto& e : h->ev)
    if (cudaEventCreate(&e) != cudaSuccess) { mpcb_destroy(h); return fail(MPCB_ERR_CUDA, "cudaEventCreate failed"); }
  for (auto& e : h->chunk_ev)
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { mpcb_destroy(h); return fail(MPCB_ERR_CUDA, "cudaEventCreate failed"); }
  if (cudaEventCreateWithFlags(&h->busy_ev, cudaEventDisableTiming) != cudaSuccess) { mpcb_destroy(h); return fail(MPCB_ERR_CUDA, "cudaEventCreate failed"); }
  rc = upload_design(h);
  if (rc != MPCB_OK) { std::string keep = g_err; mpcb_destroy(h); return fail(rc, keep); }
  if (cudaEventCreateWithFlags(&h->fan_ev, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&h->done_ev, cudaEventDisableTiming) != cudaSuccess) {
    mpcb_destroy(h); return fail(MPCB_ERR_CUDA, "cudaEventCreate failed");
  }
  // multi-device handle: the same design replicated on the other devices (each peer is an ordinary single-device handle with the same
  // kernel choice), peer access switched on both ways so that the device entry's shard copies go over NVLink directly
  for (size_t i = 1; i < dev_ids.size(); i++) {
    mpcb_settings ps = st;
    ps.n_devices = 0; ps.device = dev_ids[i]; ps.kernel = h->info.kernel;
    mpcb_handle* peer = nullptr;
    rc = mpcb_create_linear(desc, &ps, &peer);
    if (rc != MPCB_OK) { std::string keep = g_err; mpcb_destroy(h); return fail(rc, "device " + std::to_string(dev_ids[i]) + ": " + keep); }
    h->peers.push_back(peer);
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, dev_ids[i], st.device) == cudaSuccess && can) { cudaSetDevice(dev_ids[i]); cudaDeviceEnablePeerAccess(st.device, 0); }
    if (cudaDeviceCanAccessPeer(&can, st.device, dev_ids[i]) == cudaSuccess && can) { cudaSetDevice(st.device); cudaDeviceEnablePeerAccess(dev_ids[i], 0); }
    cudaGetLastError();      // "already enabled" is not an error here
  }
  cudaSetDevice(st.device);
  h->st.n_devices = (int32_t)dev_ids.size();
  *out = h;
  return MPCB_OK;
}

int mpcb_tune_rho(const mpcb_linear_desc* desc, const mpcb_settings* settings, const mpcb_batch_io* sample, int32_t n_candidates, double factor,
                  double* best_rho, double* cand_rho, double* cand_mean_iters) {
  if (!desc || !sample || !best_rho) return fail(MPCB_ERR_INVALID, "mpcb_tune_rho: null argument");
  if (n_candidates < 1 || n_candidates > 16 || !(factor > 1.0)) return fail(MPCB_ERR_INVALID, "mpcb_tune_rho: 1..16 candidates, factor > 1");
  if (sample->batch <= 0 || !sample->x0 || !sample->xref || !sample->uref) return fail(MPCB_ERR_INVALID, "mpcb_tune_rho: the sample needs x0, xref, uref");
  mpcb_settings st;
  if (settings) st = *settings; else mpcb_default_settings(&st);
  st.n_devices = 0; st.ladder_iter = 0;
  const long long n = sample->batch;
  std::vector<int32_t> status((size_t)n), iters((size_t)n);
  double rho0 = st.rho;
  double best = 0.0, best_score = 1e300;
  for (int c = -1; c < n_candidates; c++) {
    // c = -1: the automatic value (only to learn rho0); candidates j = c - (n-1)/2
    if (c == -1 && rho0 > 0.0) continue;
    mpcb_settings sc = st;
    sc.rho = (c == -1) ? 0.0 : rho0 * std::pow(factor, (double)(c - (n_candidates - 1) / 2));
    mpcb_handle* h = nullptr;
    int rc = mpcb_create_linear(desc, &sc, &h);
    if (rc != MPCB_OK) return rc;
    if (c == -1) { rho0 = h->info.rho; mpcb_destroy(h); continue; }
    mpcb_batch_io io;
    std::memset(&io, 0, sizeof(io));
    io.batch = n; io.x0 = sample->x0; io.xref = sample->xref; io.uref = sample->uref;
    io.xref_broadcast = sample->xref_broadcast; io.uref_broadcast = sample->uref_broadcast;
    io.status = status.data(); io.iters = iters.data();
    rc = mpcb_solve_linear_batch(h, &io);                 // first solve: allocations, function attributes
    if (rc == MPCB_OK) rc = mpcb_solve_linear_batch(h, &io);   // second solve: the one that is scored
    const int cap = ((h->st.max_iter + h->st.check_every - 1) / h->st.check_every) * h->st.check_every;
    const double solve_ms = (double)h->timing.solve_ms;
    mpcb_destroy(h);
    if (rc != MPCB_OK) return rc;
    // Score = what the sample costs on the GPU: the CUDA-event time of its solve (the kernels make problems wait for the slowest member of
    // their tile / slot group / row block, so neither the mean nor the maximum iteration count alone predicts it -- quadruple tank, H = 100:
    // half the automatic rho takes the mean from 102 to 58 iterations, the maximum from 110 to 220 and the kernel from 4.6 to 9.1 ms).  Samples
    // too small for the timed path fall back to the mean over groups of 32 problems of the group's maximum iteration count.
    bool all_done = true;
    double score = 0.0;
    long long groups = 0;
    for (long long g0 = 0; g0 < n; g0 += 32, groups++) {
      double mx = 0.0;
      for (long long i = g0; i < std::min<long long>(n, g0 + 32); i++) {
        const bool done = status[(size_t)i] == MPCB_STATUS_SOLVED || status[(size_t)i] == MPCB_STATUS_PRIMAL_INFEASIBLE;
        all_done = all_done && done;
        mx = std::max(mx, done ? (double)iters[(size_t)i] : 2.0 * cap);
      }
      score += mx;
    }
    score /= (double)std::max<long long>(groups, 1);
    if (solve_ms > 0.0) score = all_done ? solve_ms : solve_ms * 4.0;      // (a candidate that leaves problems at the iteration cap is not a candidate)
    if (cand_rho) cand_rho[c] = sc.rho;
    if (cand_mean_iters) cand_mean_iters[c] = score;
    if (score < best_score) { best_score = score; best = sc.rho; }
  }
  *best_rho = best;
  return MPCB_OK;
}

void mpcb_destroy(mpcb_handle* h) {
  if (!h) return;
  for (mpcb_handle* p : h->peers) mpcb_destroy(p);
  h->peers.clear();
  cudaSetDevice(h->st.device);
  if (h->fan_ev) cudaEventDestroy(h->fan_ev);
  if (h->done_ev) cudaEventDestroy(h->done_ev);
  if (h->stream) cudaStreamSynchronize(h->stream);
  h->remap.release();
  for (DevBuf<double>* b : {&h->Tfrag, &h->Cfrag, &h->Lt, &h->Lv, &h->lo, &h->hi, &h->rho, &h->rinv, &h->Tfrag2, &h->rho2, &h->rinv2, &h->A, &h->B, &h->Q, &h->R, &h->S, &h->P, &h->x0,
                            &h->xref, &h->uref, &h->warm_v, &h->warm_y, &h->v, &h->y, &h->pres, &h->dres, &h->u, &h->e_u, &h->x, &h->e_x,
                            &h->u0, &h->obj})
    b->release();
  for (DevBuf<double>* b : {&h->ric_stage, &h->ric_Lq, &h->rW, &h->rQ, &h->rD, &h->rX, &h->rXT, &h->rYO}) b->release();
  if (h->busy_ev) cudaEventDestroy(h->busy_ev);
  h->status.release(); h->iters.release(); h->counter.release();
  mpcb::stream_release(h->sc, h->sw);
  { mpcb::StreamWork none; mpcb::stream_release(h->sc2, none); }
  h->ladder_count.release();
  h->stage_in.release(); h->stage_out.release(); h->stage_int.release(); h->small_io.release(); h->small_dev.release();
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  for (auto& e : h->chunk_ev) if (e) cudaEventDestroy(e);
  if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int mpcb_get_info(const mpcb_handle* h, mpcb_info* info) {
  if (!h || !info) return fail(MPCB_ERR_INVALID, "null argument");
  *info = h->info;
  return MPCB_OK;
}

int mpcb_get_timing(const mpcb_handle* h, mpcb_timing* t) {
  if (!h || !t) return fail(MPCB_ERR_INVALID, "null argument");
  *t = h->timing;
  return MPCB_OK;
}

int mpcb_get_design(const mpcb_handle* h, double* Pc, double* Lq, double* G, double* Lb, double* T) {
  if (!h) return fail(MPCB_ERR_INVALID, "null handle");
  const mpcb::Design& D = h->D;
  if (Pc) std::memcpy(Pc, D.Pc.a.data(), sizeof(double) * D.Pc.a.size());
  if (Lq) std::memcpy(Lq, D.Lq.a.data(), sizeof(double) * D.Lq.a.size());
  if (G && D.mg) std::memcpy(G, D.G.a.data(), sizeof(double) * D.G.a.size());
  if (Lb && D.mg) std::memcpy(Lb, D.Lb.a.data(), sizeof(double) * D.Lb.a.size());
  if (T) std::memcpy(T, D.T.a.data(), sizeof(double) * D.T.a.size());
  return MPCB_OK;
}

int mpcb_solve_linear_batch_device(mpcb_handle* h, const mpcb_batch_io* io, void* cuda_stream) {
  if (!h || !io) return fail(MPCB_ERR_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(h->st.device));
  if (!h->peers.empty() && io->batch >= mpcb::MULTI_MIN_PER_DEVICE * (long long)(1 + h->peers.size())) {
    if (!io->x0 || !io->xref || !io->uref) return fail(MPCB_ERR_INVALID, "x0, xref, uref are required");
    return enqueue_device_multi(h, *io, (cudaStream_t)cuda_stream);
  }
  return enqueue_device(h, *io, (cudaStream_t)cuda_stream, nullptr);
}

int mpcb_solve_linear_batch(mpcb_handle* h, const mpcb_batch_io* hio) {
  if (!h || !hio) return fail(MPCB_ERR_INVALID, "null argument");
  const int ndev = 1 + (int)h->peers.size();
  if (ndev == 1 || hio->batch < mpcb::MULTI_MIN_PER_DEVICE * ndev) return solve_linear_host(h, hio);
  if (!hio->x0 || !hio->xref || !hio->uref) return fail(MPCB_ERR_INVALID, "x0, xref, uref are required");
  if ((hio->warm_u == nullptr) != (hio->warm_y == nullptr)) return fail(MPCB_ERR_INVALID, "warm_u and warm_y must be given together");
  // one host thread per device, each running the ordinary single-device entry on its contiguous shard of the caller's arrays
  const mpcb::ShardDims d = dims_of(h);
  int rc = mpcb::run_sharded(ndev, [&](int r) {
    long long lo, hi;
    mpcb::shard_range(hio->batch, r, ndev, &lo, &hi);
    if (hi <= lo) return (int)MPCB_OK;
    const mpcb_batch_io sio = mpcb::shard_io(*hio, lo, hi - lo, d);
    return solve_linear_host(r == 0 ? h : h->peers[(size_t)r - 1], &sio);
  });
  if (rc != MPCB_OK) return rc;
  mpcb_timing t = h->timing;          // shard 0's; the whole job: slowest shard, summed counts
  for (mpcb_handle* p : h->peers) {
    t.total_ms = std::max(t.total_ms, p->timing.total_ms); t.h2d_ms = std::max(t.h2d_ms, p->timing.h2d_ms); t.solve_ms = std::max(t.solve_ms, p->timing.solve_ms);
    t.recover_ms = std::max(t.recover_ms, p->timing.recover_ms); t.d2h_ms = std::max(t.d2h_ms, p->timing.d2h_ms);
    t.total_iterations += p->timing.total_iterations; t.kernel_launches += p->timing.kernel_launches;
  }
  t.batch = hio->batch;
  h->timing = t;
  CUDA_TRY(cudaSetDevice(h->st.device));
  return MPCB_OK;
}

}  // extern "C"

namespace {
int solve_linear_host(mpcb_handle* h, const mpcb_batch_io* hio) {
  const mpcb::Design& D = h->D;
  const long long Bn = hio->batch;
  if (Bn <= 0) return fail(MPCB_ERR_INVALID, "batch must be positive");
  if (!hio->x0 || !hio->xref || !hio->uref) return fail(MPCB_ERR_INVALID, "x0, xref, uref are required");
  if ((hio->warm_u == nullptr) != (hio->warm_y == nullptr)) return fail(MPCB_ERR_INVALID, "warm_u and warm_y must be given together");
  CUDA_TRY(cudaSetDevice(h->st.device));
  cudaStream_t st = h->stream;
  const size_t nx = D.nx, nu = D.nu, H = D.H, nz = D.nz, nt = D.nt;
  const size_t n_xref = hio->xref_broadcast ? nx : nx * Bn, n_uref = hio->uref_broadcast ? nu : nu * Bn;

  // ---- small batches (the closed-loop, one-problem-at-a-time use of update_initialization! / calculate!): no copy
  // engine at all.  Inputs are placed in one page-locked block that the kernels read over PCIe directly, the results are
  // written by the kernels straight into the same block (zero-copy), so a call is two kernel launches and one
  // synchronisation instead of ~14 DMA operations of a few hundred bytes each.
  {
    struct Seg { const double* src; double* dst; size_t n; };
    const size_t n_in[5] = {nx * (size_t)Bn, n_xref, n_uref, hio->warm_u ? nz * (size_t)Bn : 0, hio->warm_y ? nt * (size_t)Bn : 0};
    const size_t n_out[9] = {hio->u ? nz * (size_t)Bn : 0, hio->e_u ? nz * (size_t)Bn : 0, hio->x ? nx * (H + 1) * (size_t)Bn : 0,
                             hio->e_x ? nx * (H + 1) * (size_t)Bn : 0, hio->u0 ? nu * (size_t)Bn : 0, (size_t)Bn, (size_t)Bn,
                             hio->objective ? (size_t)Bn : 0, hio->y ? nt * (size_t)Bn : 0};
    size_t total = 2 * (size_t)Bn;   // status + iters as int32 pairs -> Bn doubles, rounded up
    for (size_t n : n_in) total += (n + 1) & ~(size_t)1;
    for (size_t n : n_out) total += (n + 1) & ~(size_t)1;
    if (total * sizeof(double) <= 96 * 1024) {
      CUDA_TRY(h->small_io.ensure(total));
      double* base = h->small_io.p;
      size_t off = 0;
      auto take = [&](size_t n) { double* q = base + off; off += (n + 1) & ~(size_t)1; return q; };
      const double* srcs[5] = {hio->x0, hio->xref, hio->uref, hio->warm_u, hio->warm_y};
      double* in_p[5];
      for (int i = 0; i < 5; i++) {
        in_p[i] = n_in[i] ? take(n_in[i]) : nullptr;
        if (n_in[i]) std::memcpy(in_p[i], srcs[i], n_in[i] * sizeof(double));
      }
      double* out_p[9];
      for (int i = 0; i < 9; i++) out_p[i] = n_out[i] ? take(n_out[i]) : nullptr;
      int32_t* ints = reinterpret_cast<int32_t*>(take((size_t)Bn));
      // Two modes for the block.  Zero-copy (default): the kernels read and write the page-locked block over PCIe -- no DMA operation.  Mirrored
      // (MPCB_SMALL_MIRROR, an A/B knob): the block has a twin in device memory; one copy in (the inputs lead the block), the kernels run on device
      // memory, one copy out.  Measured, one problem, C ABI p50: with the slot kernels 58.8 (zero-copy) vs 56.1 us (mirrored); with the small-batch
      // kernels of round 2 (cooperative solve + direct recover, few dependent accesses) 39.1 vs 42.8 us.
      static const bool zero_copy = std::getenv("MPCB_SMALL_MIRROR") == nullptr;
      size_t n_in_total = 0;
      for (size_t n : n_in) n_in_total += (n + 1) & ~(size_t)1;
      double* kbase = base;                  // what the kernels see
      if (!zero_copy) {
        CUDA_TRY(h->small_dev.ensure(total));
        kbase = h->small_dev.p;
        CUDA_TRY(cudaMemcpyAsync(kbase, base, n_in_total * sizeof(double), cudaMemcpyHostToDevice, st));
      }
      auto kp = [&](const void* q) { return q ? kbase + (reinterpret_cast<const double*>(q) - base) : nullptr; };
      mpcb_batch_io dio = *hio;
      dio.x0 = kp(in_p[0]); dio.xref = kp(in_p[1]); dio.uref = kp(in_p[2]); dio.warm_u = kp(in_p[3]); dio.warm_y = kp(in_p[4]);
      dio.u = kp(out_p[0]); dio.e_u = kp(out_p[1]); dio.x = kp(out_p[2]); dio.e_x = kp(out_p[3]); dio.u0 = kp(out_p[4]);
      dio.prim_res = kp(out_p[5]); dio.dual_res = kp(out_p[6]); dio.objective = kp(out_p[7]); dio.y = kp(out_p[8]);
      dio.status = reinterpret_cast<int32_t*>(kp(ints)); dio.iters = dio.status + Bn;
      int rc = enqueue_device(h, dio, st, nullptr);
      if (rc != MPCB_OK) return rc;
      if (!zero_copy) CUDA_TRY(cudaMemcpyAsync(base + n_in_total, kbase + n_in_total, (total - n_in_total) * sizeof(double), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaStreamSynchronize(st));
      double* dsts[9] = {hio->u, hio->e_u, hio->x, hio->e_x, hio->u0, hio->prim_res, hio->dual_res, hio->objective, hio->y};
      for (int i = 0; i < 9; i++)
        if (dsts[i] && n_out[i]) std::memcpy(dsts[i], out_p[i], n_out[i] * sizeof(double));
      if (hio->status) std::memcpy(hio->status, ints, Bn * sizeof(int32_t));
      if (hio->iters) std::memcpy(hio->iters, ints + Bn, Bn * sizeof(int32_t));
      long long tot = 0;
      for (long long i = 0; i < Bn; i++) tot += ints[Bn + i];
      h->timing.total_iterations = tot;
      h->timing.h2d_ms = h->timing.solve_ms = h->timing.recover_ms = h->timing.d2h_ms = h->timing.total_ms = 0.f;   // no events on this path
      h->timing.chunks = 1;
      return MPCB_OK;
    }
  }

  // ---- large batches in page-locked user memory: the result download (PCIe, ~2 KB per quadruple-tank problem) dwarfs the
  // solve, so the batch is cut into chunks and chunk c's download (copy stream) overlaps chunk c+1's upload + solve +
  // recover (compute stream).  Per-problem results do not depend on the chunking.
  {
    const double* in_ptrs[5] = {hio->x0, hio->xref, hio->uref, hio->warm_u, hio->warm_y};
    double* out_ptrs[9] = {hio->u, hio->e_u, hio->x, hio->e_x, hio->u0, hio->prim_res, hio->dual_res, hio->objective, hio->y};
    const size_t per_out[9] = {nz, nz, nx * (H + 1), nx * (H + 1), nu, 1, 1, 1, nt};
    size_t out_bytes = 0;
    bool pinned = (hio->status == nullptr || is_pinned_or_device(hio->status)) && (hio->iters == nullptr || is_pinned_or_device(hio->iters));
    const bool bc_in[5] = {false, hio->xref_broadcast != 0, hio->uref_broadcast != 0, false, false};
    for (int i = 0; i < 5; i++) if (in_ptrs[i] && !bc_in[i] && !is_pinned_or_device(in_ptrs[i])) pinned = false;   // broadcast references are tiny: staged below
    for (int i = 0; i < 9; i++)
      if (out_ptrs[i]) { out_bytes += per_out[i] * (size_t)Bn * sizeof(double); if (!is_pinned_or_device(out_ptrs[i])) pinned = false; }
    if (pinned && h->info.kernel != MPCB_KERNEL_STREAMED && out_bytes >= ((size_t)32 << 20) && Bn >= 4096) {
      int nch = (int)std::min<size_t>(8, std::max<size_t>(2, out_bytes / ((size_t)16 << 20)));   // 16 chunks measured slower (2.81 vs 2.73 ms on the bench workload)
      int first_div = 1;        // first chunk = 1/first_div of the others (its solve is the only one no download overlaps).  Measured on B200:
                                // 4..10 chunks x first_div 1..4 all land within 2.68-2.77 ms -- the 133 MB download at ~50 GB/s is the floor
      if (const char* e = std::getenv("MPCB_NCH")) { const int v = std::atoi(e); if (v >= 2 && v <= 16) nch = v; }            // experiment knobs
      if (const char* e = std::getenv("MPCB_FIRST_DIV")) { const int v = std::atoi(e); if (v >= 1 && v <= 16) first_div = v; }
      // chunk c covers [cb[c], cb[c+1]): nch - 1 equal chunks of Bc problems preceded by one of Bc / first_div
      long long cb[17];
      {
        const double units = (double)(nch - 1) + 1.0 / first_div;
        const long long Bc = (long long)std::ceil((double)Bn / units);
        long long first = std::max<long long>(1, Bc / first_div);
        cb[0] = 0;
        for (int c = 1; c <= nch; c++) cb[c] = std::min<long long>(Bn, first + (long long)(c - 1) * Bc);
        cb[nch] = Bn;
      }
      DevBuf<double>* in_dev[5] = {&h->x0, &h->xref, &h->uref, &h->warm_v, &h->warm_y};
      const size_t per_in[5] = {nx, hio->xref_broadcast ? 0 : nx, hio->uref_broadcast ? 0 : nu, nz, nt};
      const size_t n_in[5] = {nx * (size_t)Bn, n_xref, n_uref, nz * (size_t)Bn, nt * (size_t)Bn};
      for (int i = 0; i < 5; i++) if (in_ptrs[i]) CUDA_TRY(in_dev[i]->ensure(n_in[i]));
      DevBuf<double>* out_dev[9] = {&h->u, &h->e_u, &h->x, &h->e_x, &h->u0, &h->pres, &h->dres, &h->obj, &h->y};
      for (int i = 0; i < 9; i++) if (out_ptrs[i]) CUDA_TRY(out_dev[i]->ensure(per_out[i] * (size_t)Bn));
      CUDA_TRY(h->status.ensure(Bn)); CUDA_TRY(h->iters.ensure(Bn)); CUDA_TRY(h->stage_int.ensure(2 * (size_t)Bn));
      CUDA_TRY(cudaEventRecord(h->ev[0], st));
      CUDA_TRY(h->stage_in.ensure(nx + nu));
      for (int i = 1; i < 3; i++)          // broadcast references: once, through the handle's page-locked staging block
        if (per_in[i] == 0) {
          double* stg = h->stage_in.p + (i == 1 ? 0 : nx);
          std::memcpy(stg, in_ptrs[i], n_in[i] * sizeof(double));
          CUDA_TRY(cudaMemcpyAsync(in_dev[i]->p, stg, n_in[i] * sizeof(double), cudaMemcpyHostToDevice, st));
        }
      int launches = 0;
      for (int c = 0; c < nch; c++) {
        const long long b0 = cb[c], bn = cb[c + 1] - b0;
        if (bn <= 0) continue;
        for (int i = 0; i < 5; i++)
          if (in_ptrs[i] && per_in[i])
            CUDA_TRY(cudaMemcpyAsync(in_dev[i]->p + b0 * per_in[i], in_ptrs[i] + b0 * per_in[i], bn * per_in[i] * sizeof(double), cudaMemcpyHostToDevice, st));
        mpcb_batch_io dio = *hio;
        dio.batch = bn;
        dio.x0 = h->x0.p + b0 * nx; dio.xref = h->xref.p + b0 * per_in[1]; dio.uref = h->uref.p + b0 * per_in[2];
        dio.warm_u = hio->warm_u ? h->warm_v.p + b0 * nz : nullptr; dio.warm_y = hio->warm_y ? h->warm_y.p + b0 * nt : nullptr;
        double** slots[9] = {&dio.u, &dio.e_u, &dio.x, &dio.e_x, &dio.u0, &dio.prim_res, &dio.dual_res, &dio.objective, &dio.y};
        for (int i = 0; i < 9; i++) *slots[i] = out_ptrs[i] ? out_dev[i]->p + b0 * per_out[i] : nullptr;
        dio.status = h->status.p + b0; dio.iters = h->iters.p + b0;
        int rc = enqueue_device(h, dio, st, nullptr);
        if (rc != MPCB_OK) return rc;
        launches += h->timing.kernel_launches;
        CUDA_TRY(cudaEventRecord(h->chunk_ev[c], st));
        CUDA_TRY(cudaStreamWaitEvent(h->copy_stream, h->chunk_ev[c], 0));
        // per chunk only the per-problem matrices (hundreds of bytes per problem); the per-problem scalars (u0, objective,
        // residuals, status, iterations: a few bytes each) go in one copy per array after the loop -- a copy has a fixed cost
        // of a few microseconds on the copy engine, which 6 small arrays x 8 chunks would pay 48 times
        for (int i = 0; i < 9; i++)
          if (out_ptrs[i] && per_out[i] > (size_t)nu)
            CUDA_TRY(cudaMemcpyAsync(out_ptrs[i] + b0 * per_out[i], out_dev[i]->p + b0 * per_out[i], bn * per_out[i] * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream));
      }
      for (int i = 0; i < 9; i++)
        if (out_ptrs[i] && per_out[i] <= (size_t)nu)
          CUDA_TRY(cudaMemcpyAsync(out_ptrs[i], out_dev[i]->p, (size_t)Bn * per_out[i] * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream));
      CUDA_TRY(cudaMemcpyAsync(h->stage_int.p, h->status.p, Bn * sizeof(int32_t), cudaMemcpyDeviceToHost, h->copy_stream));
      CUDA_TRY(cudaMemcpyAsync(h->stage_int.p + Bn, h->iters.p, Bn * sizeof(int32_t), cudaMemcpyDeviceToHost, h->copy_stream));
      CUDA_TRY(cudaEventRecord(h->ev[4], h->copy_stream));
      CUDA_TRY(cudaStreamSynchronize(h->copy_stream));
      CUDA_TRY(cudaStreamSynchronize(st));
      if (hio->status) std::memcpy(hio->status, h->stage_int.p, Bn * sizeof(int32_t));
      if (hio->iters) std::memcpy(hio->iters, h->stage_int.p + Bn, Bn * sizeof(int32_t));
      long long tot = 0;
      for (long long i = 0; i < Bn; i++) tot += h->stage_int.p[Bn + i];
      h->timing.total_iterations = tot; h->timing.batch = Bn; h->timing.kernel_launches = launches; h->timing.chunks = nch;
      h->timing.h2d_ms = h->timing.solve_ms = h->timing.recover_ms = h->timing.d2h_ms = 0.f;   // the phases overlap: only the total is meaningful
      cudaEventElapsedTime(&h->timing.total_ms, h->ev[0], h->ev[4]);
      return MPCB_OK;
    }
  }
  h->timing.chunks = 1;

  // ---- inputs: pageable user memory goes through one pinned staging buffer, pinned memory is copied directly
  struct In { const double* src; DevBuf<double>* dst; size_t n; };
  In ins[5] = {{hio->x0, &h->x0, nx * Bn}, {hio->xref, &h->xref, n_xref}, {hio->uref, &h->uref, n_uref},
               {hio->warm_u, &h->warm_v, nz * Bn}, {hio->warm_y, &h->warm_y, nt * Bn}};
  size_t stage_need = 0;
  for (auto& in : ins)
    if (in.src && !is_pinned_or_device(in.src)) stage_need += in.n;
  CUDA_TRY(h->stage_in.ensure(stage_need));
  CUDA_TRY(cudaEventRecord(h->ev[0], st));
  size_t off = 0;
  for (auto& in : ins) {
    if (!in.src) continue;
    CUDA_TRY(in.dst->ensure(in.n));
    const double* src = in.src;
    if (!is_pinned_or_device(in.src)) {
      std::memcpy(h->stage_in.p + off, in.src, in.n * sizeof(double));
      src = h->stage_in.p + off;
      off += in.n;
    }
    CUDA_TRY(cudaMemcpyAsync(in.dst->p, src, in.n * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  CUDA_TRY(cudaEventRecord(h->ev[1], st));

  // ---- device buffers for the requested outputs
  mpcb_batch_io dio = *hio;
  dio.x0 = h->x0.p; dio.xref = h->xref.p; dio.uref = h->uref.p;
  dio.warm_u = hio->warm_u ? h->warm_v.p : nullptr; dio.warm_y = hio->warm_y ? h->warm_y.p : nullptr;
  struct Out { double* host; DevBuf<double>* dev; size_t n; double** slot; };
  Out outs[9] = {{hio->u, &h->u, nu * H * Bn, &dio.u}, {hio->e_u, &h->e_u, nu * H * Bn, &dio.e_u}, {hio->x, &h->x, nx * (H + 1) * Bn, &dio.x},
                 {hio->e_x, &h->e_x, nx * (H + 1) * Bn, &dio.e_x}, {hio->u0, &h->u0, nu * Bn, &dio.u0}, {hio->prim_res, &h->pres, (size_t)Bn, &dio.prim_res},
                 {hio->dual_res, &h->dres, (size_t)Bn, &dio.dual_res}, {hio->objective, &h->obj, (size_t)Bn, &dio.objective}, {hio->y, &h->y, nt * Bn, &dio.y}};
  for (auto& o : outs) {
    if (!o.host) { *o.slot = nullptr; continue; }
    CUDA_TRY(o.dev->ensure(o.n));
    *o.slot = o.dev->p;
  }
  CUDA_TRY(h->status.ensure(Bn)); CUDA_TRY(h->iters.ensure(Bn));
  dio.status = h->status.p; dio.iters = h->iters.p;

  int rc = enqueue_device(h, dio, st, h->ev[2]);
  if (rc != MPCB_OK) return rc;
  CUDA_TRY(cudaEventRecord(h->ev[3], st));

  // ---- outputs
  size_t out_stage = 0;
  for (auto& o : outs)
    if (o.host && !is_pinned_or_device(o.host)) out_stage += o.n;
  CUDA_TRY(h->stage_out.ensure(out_stage));
  CUDA_TRY(h->stage_int.ensure(2 * (size_t)Bn));
  off = 0;
  std::vector<std::pair<double*, std::pair<size_t, size_t>>> post;  // host dst, (stage offset, n)
  for (auto& o : outs) {
    if (!o.host) continue;
    if (is_pinned_or_device(o.host)) {
      CUDA_TRY(cudaMemcpyAsync(o.host, o.dev->p, o.n * sizeof(double), cudaMemcpyDeviceToHost, st));
    } else {
      CUDA_TRY(cudaMemcpyAsync(h->stage_out.p + off, o.dev->p, o.n * sizeof(double), cudaMemcpyDeviceToHost, st));
      post.push_back({o.host, {off, o.n}});
      off += o.n;
    }
  }
  CUDA_TRY(cudaMemcpyAsync(h->stage_int.p, h->status.p, Bn * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(h->stage_int.p + Bn, h->iters.p, Bn * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaEventRecord(h->ev[4], st));
  CUDA_TRY(cudaStreamSynchronize(st));
  for (auto& p : post) std::memcpy(p.first, h->stage_out.p + p.second.first, p.second.second * sizeof(double));
  if (hio->status) std::memcpy(hio->status, h->stage_int.p, Bn * sizeof(int32_t));
  if (hio->iters) std::memcpy(hio->iters, h->stage_int.p + Bn, Bn * sizeof(int32_t));
  long long tot = 0;
  for (long long i = 0; i < Bn; i++) tot += h->stage_int.p[Bn + i];
  h->timing.total_iterations = tot;
  cudaEventElapsedTime(&h->timing.h2d_ms, h->ev[0], h->ev[1]);
  cudaEventElapsedTime(&h->timing.solve_ms, h->ev[1], h->ev[2]);
  cudaEventElapsedTime(&h->timing.recover_ms, h->ev[2], h->ev[3]);
  cudaEventElapsedTime(&h->timing.d2h_ms, h->ev[3], h->ev[4]);
  cudaEventElapsedTime(&h->timing.total_ms, h->ev[0], h->ev[4]);
  return MPCB_OK;
}

int closed_loop_linear_single(mpcb_handle* h, const mpcb_closed_loop_io* cio);
}  // namespace

extern "C" {

int mpcb_closed_loop_linear_batch(mpcb_handle* h, const mpcb_closed_loop_io* cio) {
  if (!h || !cio) return fail(MPCB_ERR_INVALID, "null argument");
  const int ndev = 1 + (int)h->peers.size();
  if (ndev == 1 || cio->batch < mpcb::MULTI_MIN_PER_DEVICE * ndev) return closed_loop_linear_single(h, cio);
  if (!cio->x0 || !cio->xref || !cio->uref) return fail(MPCB_ERR_INVALID, "x0, xref, uref are required");
  int rc = mpcb::run_sharded(ndev, [&](int r) {
    long long lo, hi;
    mpcb::shard_range(cio->batch, r, ndev, &lo, &hi);
    if (hi <= lo) return (int)MPCB_OK;
    const mpcb_closed_loop_io sio = mpcb::shard_closed_loop_io(*cio, lo, hi - lo, (size_t)h->D.nx, (size_t)h->D.nu);
    return closed_loop_linear_single(r == 0 ? h : h->peers[(size_t)r - 1], &sio);
  });
  if (rc != MPCB_OK) return rc;
  for (mpcb_handle* p : h->peers) { h->timing.total_ms = std::max(h->timing.total_ms, p->timing.total_ms); h->timing.kernel_launches += p->timing.kernel_launches; }
  h->timing.batch = cio->batch;
  CUDA_TRY(cudaSetDevice(h->st.device));
  return MPCB_OK;
}

}  // extern "C"

namespace {
int closed_loop_linear_single(mpcb_handle* h, const mpcb_closed_loop_io* cio) {
  const mpcb::Design& D = h->D;
  const long long Bn = cio->batch;
  const int T = cio->steps;
  if (Bn <= 0 || T <= 0) return fail(MPCB_ERR_INVALID, "batch and steps must be positive");
  if (!cio->x0 || !cio->xref || !cio->uref) return fail(MPCB_ERR_INVALID, "x0, xref, uref are required");
  CUDA_TRY(cudaSetDevice(h->st.device));
  cudaStream_t st = h->stream;
  const size_t nx = D.nx, nu = D.nu, nz = D.nz, nt = D.nt, B = (size_t)Bn;
  const size_t n_xref = cio->xref_broadcast ? nx : nx * B, n_uref = cio->uref_broadcast ? nu : nu * B;
  // state ping-pong in x0 / warm_v's neighbours: dedicated buffers keep this independent of the batch entry's workspaces
  DevBuf<double> xa, xb, va, vb, ya, yb, xtraj, utraj;
  DevBuf<int32_t> itot, unsol;
  auto cleanup = [&]() { xa.release(); xb.release(); va.release(); vb.release(); ya.release(); yb.release(); xtraj.release(); utraj.release(); itot.release(); unsol.release(); };
#define CL_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); return fail(MPCB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } } while (0)
  CL_TRY(xa.ensure(nx * B)); CL_TRY(xb.ensure(nx * B)); CL_TRY(va.ensure(nz * B)); CL_TRY(vb.ensure(nz * B)); CL_TRY(ya.ensure(nt * B)); CL_TRY(yb.ensure(nt * B));
  CL_TRY(h->xref.ensure(n_xref)); CL_TRY(h->uref.ensure(n_uref)); CL_TRY(h->status.ensure(B)); CL_TRY(h->iters.ensure(B));
  if (cio->x_traj) CL_TRY(xtraj.ensure(nx * (size_t)(T + 1) * B));
  if (cio->u_traj) CL_TRY(utraj.ensure(nu * (size_t)T * B));
  if (cio->iters_total) CL_TRY(itot.ensure(B));
  if (cio->unsolved_steps) CL_TRY(unsol.ensure(B));
  CL_TRY(cudaMemcpyAsync(xa.p, cio->x0, nx * B * sizeof(double), cudaMemcpyHostToDevice, st));
  CL_TRY(cudaMemcpyAsync(h->xref.p, cio->xref, n_xref * sizeof(double), cudaMemcpyHostToDevice, st));
  CL_TRY(cudaMemcpyAsync(h->uref.p, cio->uref, n_uref * sizeof(double), cudaMemcpyHostToDevice, st));
  CL_TRY(cudaEventRecord(h->ev[0], st));
  int launches = 0;
  double *xc = xa.p, *xn = xb.p, *vc = va.p, *vp = vb.p, *yc = ya.p, *yp = yb.p;
  for (int t = 0; t < T; t++) {
    mpcb_batch_io dio;
    std::memset(&dio, 0, sizeof(dio));
    dio.batch = Bn; dio.x0 = xc; dio.xref = h->xref.p; dio.uref = h->uref.p; dio.xref_broadcast = cio->xref_broadcast; dio.uref_broadcast = cio->uref_broadcast;
    if (cio->warm_start && t > 0) { dio.warm_u = vp; dio.warm_y = yp; }
    dio.status = h->status.p; dio.iters = h->iters.p;
    dio.y = cio->warm_start ? yc : nullptr;
    int rc = enqueue_device(h, dio, st, nullptr, vc);      // the absolute inputs land in vc: u0 for the plant, warm start for t+1
    if (rc != MPCB_OK) { cleanup(); return rc; }
    launches += h->timing.kernel_launches;
    mpcb::PlantStepParams Pp;
    Pp.A = h->A.p; Pp.B = h->B.p; Pp.nx = D.nx; Pp.nu = D.nu; Pp.nz = D.nz; Pp.t = t; Pp.steps = T; Pp.batch = Bn;
    Pp.xref = h->xref.p; Pp.uref = h->uref.p; Pp.xref_bc = cio->xref_broadcast; Pp.uref_bc = cio->uref_broadcast;
    Pp.v = vc; Pp.status = h->status.p; Pp.iters = h->iters.p; Pp.x_in = xc; Pp.x_out = xn;
    Pp.x_traj = cio->x_traj ? xtraj.p : nullptr; Pp.u_traj = cio->u_traj ? utraj.p : nullptr;
    Pp.iters_total = cio->iters_total ? itot.p : nullptr; Pp.unsolved = cio->unsolved_steps ? unsol.p : nullptr;
    mpcb::plant_step_kernel<<<(unsigned)((Bn + 127) / 128), 128, 0, st>>>(Pp);
    CL_TRY(cudaGetLastError());
    launches += 1;
    std::swap(xc, xn); std::swap(vc, vp); std::swap(yc, yp);
  }
  CL_TRY(cudaEventRecord(h->ev[3], st));
  if (cio->x_traj) CL_TRY(cudaMemcpyAsync(cio->x_traj, xtraj.p, nx * (size_t)(T + 1) * B * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (cio->u_traj) CL_TRY(cudaMemcpyAsync(cio->u_traj, utraj.p, nu * (size_t)T * B * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (cio->iters_total) CL_TRY(cudaMemcpyAsync(cio->iters_total, itot.p, B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (cio->unsolved_steps) CL_TRY(cudaMemcpyAsync(cio->unsolved_steps, unsol.p, B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CL_TRY(cudaEventRecord(h->ev[4], st));
  CL_TRY(cudaStreamSynchronize(st));
#undef CL_TRY
  cudaEventElapsedTime(&h->timing.solve_ms, h->ev[0], h->ev[3]);
  cudaEventElapsedTime(&h->timing.d2h_ms, h->ev[3], h->ev[4]);
  h->timing.h2d_ms = 0.f; h->timing.recover_ms = 0.f; h->timing.total_ms = h->timing.solve_ms + h->timing.d2h_ms;
  h->timing.batch = Bn; h->timing.kernel_launches = launches; h->timing.chunks = 1; h->timing.total_iterations = 0;
  cleanup();
  return MPCB_OK;
}

}  // namespace
