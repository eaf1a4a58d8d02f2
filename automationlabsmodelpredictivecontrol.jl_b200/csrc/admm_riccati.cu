// Stage-wise ("sparse" / Riccati) batched ADMM for the linear-MPC QP, sm_100a.
//
// The reference's own formulation of the QP is the sparse, stage-wise one -- variables linear in the horizon
// (/root/reference/src/sub/model_modeler_implementation/linear/mpc_modeler_implementation_linear.jl:48-60).  The condensed
// kernels (admm_onchip.cuh, admm_smem.cuh, admm_stream.cu) trade that structure for one dense nz x nz contraction per
// iteration: 2 nz^2 = 8 H^2 flops (quadruple tank) on the FP64 tensor pipe.  This kernel keeps the stage structure.  It runs
// the SAME ADMM iteration (same rho, sigma, alpha, same iterates up to round-off, same termination rule -- see
// admm_onchip.cuh) but solves the x-update
//     K x~ = r,   K = Pc + (sigma + rho) I        (box-only problems)
// as what it is: the optimality system of an unconstrained finite-horizon LQ problem with zero initial state, modified input
// weight R^ = 2R + (sigma + rho) I and the linear input cost -r.  Its solution is one backward and one forward sweep with
// matrices that depend on the system only and are cached per stage at design time (host_design.cpp, riccati_factors):
//     backward, k = H-1 .. 0:   s = r_k + B' pi;   d_k = Lam_k^-1 s;   pi <- Acl_k' pi - K_k' r_k
//     forward,  k = 0 .. H-1:   u_k = d_k - K_k e;  e <- Acl_k e + B d_k           (x~ = [u_0 .. u_{H-1}])
// i.e. 2 nx^2 + 4 nx nu + nu^2 multiply-adds per stage (68 for the quadruple tank) instead of 2 nu^2 H: the crossover in
// arithmetic is at H ~ 17, and from there on the kernel is bound by streaming its per-problem state, not by FP64 issue.
//
// Mapping: one lane = one problem, one warp = one TILE of 32 problems.  All per-problem state lives in HBM/L2 in a
// tile-major layout [tile][row][lane]: a row of a tile is 256 contiguous bytes, a chunk of `ch` stages is one contiguous
// block.  State per row: only w = alpha x~ + (1-alpha) z + y/rho of the previous iteration (z = clamp(w), y/rho = w - z are
// recomputed), the linear term q, the feed-forward d between the two sweeps (and the relaxed x when sigma != 0) --
// 6 row passes per iteration.  Chunks are staged into shared memory with 1-D TMA bulk copies (cp.async.bulk +
// mbarrier complete_tx), NBUF deep, issued by one lane ahead of the sweep, so that the per-warp dependent chain never waits
// on L2/HBM latency; results go back with plain stores (a row is one fully coalesced 256-byte store per warp).  The two sweeps
// traverse the horizon in opposite directions, so the chunks written last are read first and tend to hit in L2.
// Stage matrices sit in shared memory (broadcast reads); B and B' are kernel parameters, i.e. constant-bank operands.
#include "admm_riccati.cuh"

#include <algorithm>

namespace mpcb {

namespace {

constexpr int NBUF = 2;      // double buffering: one chunk (>= 4 stages x ~500 cycles of sweep) covers the L2 / HBM latency of the next; the third buffer of the first version only cost chunk length
constexpr int WARP_HDR = 16;   // doubles reserved per warp in front of its buffers (mbarriers), keeps the buffers 128-byte aligned

constexpr int ev2(int n) { return (n + 1) & ~1; }
template <int NX, int NU>
struct SL {   // per-stage block, offsets in doubles; every matrix starts on a 16-byte boundary
  static constexpr int K = 0;                        // [NU][NX]
  static constexpr int KT = K + ev2(NU * NX);        // [NX][NU]
  static constexpr int LI = KT + ev2(NX * NU);       // [NU][NU]
  static constexpr int ACL = LI + ev2(NU * NU);      // [NX][NX]
  static constexpr int ACLT = ACL + ev2(NX * NX);    // [NX][NX]
  static constexpr int SIZE = ACLT + ev2(NX * NX);
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "RIC_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra RIC_DONE;\n"
      "bra RIC_WAIT;\n"
      "RIC_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
// generic-proxy writes (plain stores) before async-proxy reads (bulk copies) of the same memory
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ double dmaxf(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double dclamp(double w, double lo, double hi) { const double t = w < lo ? lo : w; return t > hi ? hi : t; }

template <int NX, int NU, bool SIG>
__global__ void __launch_bounds__(256) admm_riccati_kernel(const __grid_constant__ RiccatiParams P) {
  using L = SL<NX, NU>;
  constexpr int NP = 2 * NX + NU;
  constexpr int NARR = SIG ? 4 : 3;                       // staged input arrays at once: W, Qb, D (+ X)
  extern __shared__ __align__(128) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int H = P.H, nz = P.nz, ch = P.ch;
  const int stage_doubles = ((H * L::SIZE + 15) / 16) * 16;
  const int chunk_doubles = ch * NU * 32;
  double* sStage = smem;
  double* wbase = smem + stage_doubles + (size_t)warp * (WARP_HDR + NARR * NBUF * chunk_doubles);
  uint64_t* bar = reinterpret_cast<uint64_t*>(wbase);
  double* buf0 = wbase + WARP_HDR;
  auto buf = [&](int arr, int b) { return buf0 + (size_t)(arr * NBUF + b) * chunk_doubles; };

  for (int i = threadIdx.x; i < H * L::SIZE; i += blockDim.x) sStage[i] = P.stage[i];
  if (lane == 0) {
    for (int b = 0; b < NBUF; b++) mbar_init(bar + b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const double rho = P.rho, sigma = P.sigma, alpha = P.alpha, oma = 1.0 - P.alpha, sig_rho = P.sigma + P.rho;
  const long long ntiles = (P.batch + 31) / 32;
  const int NC = (H + ch - 1) / ch;
  const int max_iter = ((P.max_iter + P.check_every - 1) / P.check_every) * P.check_every;
  uint32_t phase = 0;   // bit b: parity the next wait on buffer b expects
  uint32_t gs = 0;      // running step counter: step gs uses buffer gs % NBUF

  for (;;) {
    long long tile = 0;
    if (lane == 0) tile = (long long)atomicAdd(P.counter, 1ULL);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile >= ntiles) break;
    const bool valid = tile * 32 + lane < P.batch;
    const long long p = valid ? tile * 32 + lane : P.batch - 1;
    const size_t toff = (size_t)tile * nz * 32;
    double* Wt = P.W + toff; double* Qt = P.Qb + toff; double* Dt = P.D + toff; double* XTt = P.XT + toff;
    double* Xt = SIG ? P.X + toff : nullptr;
    double* YOt = P.YO ? P.YO + toff : nullptr;

    // ---------------------------------------------------------------- setup: q = Lq p, |q|_inf, warm start
    double qn = 0.0;
    {
      double pv[NP];
#pragma unroll
      for (int j = 0; j < NP; j++) {
        if (j < NX) pv[j] = P.x0[p * NX + j];
        else if (j < 2 * NX) pv[j] = P.xref[(P.xref_bc ? 0 : p) * NX + (j - NX)];
        else pv[j] = P.uref[(P.uref_bc ? 0 : p) * NU + (j - 2 * NX)];
      }
      for (int j = 0; j < nz; j++) {
        double q = 0.0;
#pragma unroll
        for (int i = 0; i < NP; i++) q = fma(__ldg(P.Lq + (size_t)j * NP + i), pv[i], q);
        Qt[j * 32 + lane] = q;
        qn = dmaxf(qn, fabs(q));
        if (P.warm_v != nullptr) {
          // warm start (x, y) -> w = x + y/rho: z = clamp(w), y/rho = w - z  (for a KKT pair exactly (x, y) again)
          const double v0 = P.warm_v[p * nz + j], y0 = P.warm_y[p * nz + j];
          Wt[j * 32 + lane] = v0 + y0 / rho;
          if (SIG) Xt[j * 32 + lane] = v0;
        } else if (P.Lv != nullptr) {
          // settings.cold_init: x = clip(v_unc), y / rho = -kappa (x - v_unc) with v_unc = Lv p the unconstrained optimum
          double vu = 0.0;
#pragma unroll
          for (int i = 0; i < NP; i++) vu = fma(__ldg(P.Lv + (size_t)j * NP + i), pv[i], vu);
          const double v0 = dclamp(vu, P.lo[j % NU], P.hi[j % NU]);
          Wt[j * 32 + lane] = v0 - MPCB_INIT_KAPPA * (v0 - vu);
          if (SIG) Xt[j * 32 + lane] = v0;
        }
      }
    }
    fence_proxy_async();
    __syncwarp();

    bool done = !valid;
    bool cold_first = (P.warm_v == nullptr && P.Lv == nullptr);   // first iteration of a cold start: z = y = x = 0, nothing to read but q
    int it = 0;

    // one chunk of one array set: lane 0 arms the barrier and issues the bulk copies
    auto issue = [&](int c, int b, bool ldW, bool ldQ, bool ldD, bool ldX) {
      if (lane == 0) {
        const int k0 = c * ch, k1 = min(H, k0 + ch);
        const uint32_t bytes = (uint32_t)((k1 - k0) * NU * 32 * sizeof(double));
        const size_t off = (size_t)k0 * NU * 32;
        const int n = (ldW ? 1 : 0) + (ldQ ? 1 : 0) + (ldD ? 1 : 0) + ((SIG && ldX) ? 1 : 0);
        mbar_expect_tx(bar + b, bytes * n);
        if (ldW) bulk_g2s(buf(0, b), Wt + off, bytes, bar + b);
        if (ldQ) bulk_g2s(buf(1, b), Qt + off, bytes, bar + b);
        if (ldD) bulk_g2s(buf(2, b), Dt + off, bytes, bar + b);
        if (SIG && ldX) bulk_g2s(buf(3, b), Xt + off, bytes, bar + b);
      }
    };

    for (;;) {
      // ================================================================ backward sweep: d_k for every stage
      {
        const bool ldS = !cold_first;   // previous state needed
        double pi[NX];
#pragma unroll
        for (int m = 0; m < NX; m++) pi[m] = 0.0;
        for (int s = 0; s < NBUF - 1 && s < NC; s++) issue(NC - 1 - s, (gs + s) % NBUF, ldS, true, false, ldS);
        for (int s = 0; s < NC; s++) {
          const int c = NC - 1 - s, b = (gs + s) % NBUF;
          if (s + NBUF - 1 < NC) {
            __syncwarp();   // every lane is done with the buffer of step s-1
            issue(c - (NBUF - 1), (gs + s + NBUF - 1) % NBUF, ldS, true, false, ldS);
          }
          mbar_wait(bar + b, (phase >> b) & 1u);
          phase ^= 1u << b;
          const int k0 = c * ch, k1 = min(H, k0 + ch);
          const double* bW = buf(0, b) + lane; const double* bQ = buf(1, b) + lane; const double* bX = buf(SIG ? 3 : 0, b) + lane;
          for (int k = k1 - 1; k >= k0; k--) {
            const double* M = sStage + (size_t)k * L::SIZE;
            double h[NU], sv[NU];
#pragma unroll
            for (int i = 0; i < NU; i++) {
              const int row = (k - k0) * NU + i;
              const double q = bQ[row * 32];
              if (cold_first) h[i] = -q;
              else {
                const double wp = bW[row * 32];
                const double zp = dclamp(wp, P.lo[i], P.hi[i]);
                h[i] = fma(rho, fma(2.0, zp, -wp), -q);
                if (SIG) h[i] = fma(sigma, bX[row * 32], h[i]);
              }
              double acc = h[i];
#pragma unroll
              for (int m = 0; m < NX; m++) acc = fma(P.Bt[i * NX + m], pi[m], acc);
              sv[i] = acc;
            }
#pragma unroll
            for (int i = 0; i < NU; i++) {
              double d = 0.0;
#pragma unroll
              for (int l = 0; l < NU; l++) d = fma(M[L::LI + i * NU + l], sv[l], d);
              Dt[(size_t)(k * NU + i) * 32 + lane] = d;
            }
            double pn[NX];
#pragma unroll
            for (int m = 0; m < NX; m++) {
              double a0 = 0.0, a1 = 0.0;   // two partial sums: halves the dependent chain
#pragma unroll
              for (int n = 0; n < NX; n++) {
                if (n & 1) a1 = fma(M[L::ACLT + m * NX + n], pi[n], a1);
                else a0 = fma(M[L::ACLT + m * NX + n], pi[n], a0);
              }
#pragma unroll
              for (int i = 0; i < NU; i++) a1 = fma(-M[L::KT + m * NU + i], h[i], a1);
              pn[m] = a0 + a1;
            }
#pragma unroll
            for (int m = 0; m < NX; m++) pi[m] = pn[m];
          }
        }
        gs += NC;
        fence_proxy_async();
        __syncwarp();
      }
      it++;
      const bool chk = (it % P.check_every) == 0;

      // ================================================================ forward sweep: x~, then the ADMM step per row
      double rp = 0.0, rd = 0.0, nA = 0.0, nD = 0.0;
      {
        const bool ldS = !cold_first;
        double e[NX];
#pragma unroll
        for (int m = 0; m < NX; m++) e[m] = 0.0;
        for (int s = 0; s < NBUF - 1 && s < NC; s++) issue(s, (gs + s) % NBUF, ldS, chk, true, ldS);
        for (int s = 0; s < NC; s++) {
          const int c = s, b = (gs + s) % NBUF;
          if (s + NBUF - 1 < NC) {
            __syncwarp();
            issue(c + NBUF - 1, (gs + s + NBUF - 1) % NBUF, ldS, chk, true, ldS);
          }
          mbar_wait(bar + b, (phase >> b) & 1u);
          phase ^= 1u << b;
          const int k0 = c * ch, k1 = min(H, k0 + ch);
          const double* bW = buf(0, b) + lane; const double* bQ = buf(1, b) + lane; const double* bD = buf(2, b) + lane;
          const double* bX = buf(SIG ? 3 : 0, b) + lane;
          for (int k = k0; k < k1; k++) {
            const double* M = sStage + (size_t)k * L::SIZE;
            double d[NU], u[NU];
#pragma unroll
            for (int i = 0; i < NU; i++) {
              d[i] = bD[((k - k0) * NU + i) * 32];
              double a = d[i];
#pragma unroll
              for (int m = 0; m < NX; m++) a = fma(-M[L::K + i * NX + m], e[m], a);
              u[i] = a;
            }
            double en[NX];
#pragma unroll
            for (int m = 0; m < NX; m++) {
              double a0 = 0.0, a1 = 0.0;
#pragma unroll
              for (int n = 0; n < NX; n++) {
                if (n & 1) a1 = fma(M[L::ACL + m * NX + n], e[n], a1);
                else a0 = fma(M[L::ACL + m * NX + n], e[n], a0);
              }
#pragma unroll
              for (int i = 0; i < NU; i++) a1 = fma(P.Bm[m * NU + i], d[i], a1);
              en[m] = a0 + a1;
            }
#pragma unroll
            for (int m = 0; m < NX; m++) e[m] = en[m];
            // ---- ADMM step on the rows of u_k (OSQP update_x / update_z / update_y, as in admm_onchip.cuh)
#pragma unroll
            for (int i = 0; i < NU; i++) {
              const int row = (k - k0) * NU + i;
              const size_t g = (size_t)(k * NU + i) * 32 + lane;
              const double t = u[i];
              double cc = 0.0, xp = 0.0, r_old = 0.0;
              if (!cold_first) {
                const double wp = bW[row * 32];
                const double zp = dclamp(wp, P.lo[i], P.hi[i]);
                cc = fma(-alpha, zp, wp);
                if (SIG) xp = bX[row * 32];
                if (chk) { r_old = fma(rho, fma(2.0, zp, -wp), -bQ[row * 32]); if (SIG) r_old = fma(sigma, xp, r_old); }
              } else if (chk) {
                r_old = -bQ[row * 32];
              }
              const double w = fma(alpha, t, cc);
              Wt[g] = w;
              if (SIG) Xt[g] = fma(alpha, t, oma * xp);
              if (chk) {   // residuals of (x~, z+, y+): Pc x~ = r - (sigma + rho) x~
                const double zn = dclamp(w, P.lo[i], P.hi[i]);
                const double pc = fma(-sig_rho, t, r_old);
                const double yb = rho * (w - zn);
                rp = dmaxf(rp, fabs(t - zn));
                rd = dmaxf(rd, fabs(pc + bQ[row * 32] + yb));
                nA = dmaxf(nA, dmaxf(fabs(t), fabs(zn)));
                nD = dmaxf(nD, dmaxf(fabs(pc), fabs(yb)));
                if (!done) { XTt[g] = t; if (YOt) YOt[g] = yb; }
              }
            }
          }
        }
        gs += NC;
        fence_proxy_async();
        __syncwarp();
      }
      cold_first = false;
      if (chk) {
        const bool conv = (rp <= P.eps_abs + P.eps_rel * nA) && (rd <= P.eps_abs + P.eps_rel * dmaxf(nD, qn));
        if (!done && (conv || it >= max_iter)) {
          P.status[p] = conv ? 1 : -2;
          P.iters[p] = it;
          P.pres[p] = rp;
          P.dres[p] = rd;
          done = true;
        }
        if (__all_sync(0xffffffffu, done)) break;
      }
    }

    // ---------------------------------------------------------------- results of this tile, problem-major for the callers
    if (valid) {
      for (int j = 0; j < nz; j++) P.v_out[p * nz + j] = XTt[j * 32 + lane];
      if (P.y_out != nullptr)
        for (int j = 0; j < nz; j++) P.y_out[p * nz + j] = YOt[j * 32 + lane];
    }
  }

  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long prev = atomicAdd(P.counter + 1, 1ULL);
    if (prev == (unsigned long long)gridDim.x - 1ULL) { P.counter[0] = 0ULL; P.counter[1] = 0ULL; }
  }
}

template <int NX, int NU>
size_t smem_for(int H, int ch, int wpc, bool sig) {
  const size_t stage = (((size_t)H * SL<NX, NU>::SIZE + 15) / 16) * 16;
  const size_t per_warp = WARP_HDR + (size_t)(sig ? 4 : 3) * NBUF * ch * NU * 32;
  return sizeof(double) * (stage + wpc * per_warp);
}

template <int NX, int NU>
bool plan_t(int H, bool sig, long long batch, int sm_count, size_t limit, int* wpc_out, int* ch_out, size_t* smem_out) {
  const long long tiles = (batch + 31) / 32;
  int wps = (int)std::min<long long>(16, std::max<long long>(1, (tiles + sm_count - 1) / sm_count));   // warps per SM worth having
  for (; wps >= 1; wps--) {
    const int wpc = std::min(wps, 8), ctas = (wps + wpc - 1) / wpc;
    const size_t budget = (size_t)(228 * 1024) / ctas - 1024;      // per-CTA share of the SM's shared memory (1 KB reserved per CTA)
    for (int ch : {16, 8, 4, 2, 1}) {
      if (ch > H && ch > 1) continue;
      const size_t need = smem_for<NX, NU>(H, ch, wpc, sig);
      if (need <= std::min(budget, limit)) { *wpc_out = wpc; *ch_out = ch; *smem_out = need; return true; }
    }
  }
  return false;
}

template <int NX, int NU, bool SIG>
cudaError_t launch_t(RiccatiParams P, int sm_count, size_t limit, cudaStream_t st) {
  int wpc = 0, ch = 0; size_t smem = 0;
  if (!plan_t<NX, NU>(P.H, SIG, P.batch, sm_count, limit, &wpc, &ch, &smem)) return cudaErrorInvalidConfiguration;
  P.ch = ch;
  auto kern = admm_riccati_kernel<NX, NU, SIG>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, wpc * 32, smem);
  if (e != cudaSuccess) return e;
  const long long tiles = (P.batch + 31) / 32;
  const long long grid = std::min<long long>((tiles + wpc - 1) / wpc, (long long)sm_count * std::max(occ, 1));
  kern<<<(unsigned)std::max<long long>(grid, 1), wpc * 32, smem, st>>>(P);
  return cudaGetLastError();
}

}  // namespace

int ric_stage_doubles(int nx, int nu) { return ev2(nu * nx) * 2 + ev2(nu * nu) + ev2(nx * nx) * 2; }

// (nx, nu) pairs with a compiled instance
#define MPCB_RIC_SIZES(X) X(2, 1) X(3, 1) X(3, 2) X(4, 1) X(4, 2) X(5, 3) X(6, 2) X(6, 3) X(8, 4)

bool riccati_supported(int nx, int nu) {
#define X(a, b) if (nx == a && nu == b) return true;
  MPCB_RIC_SIZES(X)
#undef X
  return false;
}

bool riccati_plan(int nx, int nu, int H, bool sig, long long batch, int sm_count, size_t smem_limit, int* wpc, int* ch, size_t* smem) {
#define X(a, b) if (nx == a && nu == b) return plan_t<a, b>(H, sig, batch, sm_count, smem_limit, wpc, ch, smem);
  MPCB_RIC_SIZES(X)
#undef X
  return false;
}

cudaError_t launch_riccati(int nx, int nu, bool sig, RiccatiParams P, int sm_count, size_t smem_limit, cudaStream_t st) {
#define X(a, b) if (nx == a && nu == b) return sig ? launch_t<a, b, true>(P, sm_count, smem_limit, st) : launch_t<a, b, false>(P, sm_count, smem_limit, st);
  MPCB_RIC_SIZES(X)
#undef X
  return cudaErrorInvalidValue;
}

}  // namespace mpcb
