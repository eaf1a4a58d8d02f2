// Shared-memory resident batched ADMM for box-only problems with 64 < nt <= ~120 (sm_100a): the middle regime between
// the register-resident kernel (admm_onchip.cuh, nt <= 64) and the streamed GEMM (admm_stream.cu).
//
// Same algorithm, slot scheme and DMMA operand trick as admm_onchip.cuh (one warp = 8 problem slots, problems along the M
// dimension of DMMA.8x8x4, T in host-prepared fragment order as the B operand, the C fragment of one iteration already in
// the A-fragment layout of the next), but the per-row state c, r, q (and x when sigma > 0) does not fit the register file
// any more: it lives in a per-warp shared-memory slice in [row][lane] order (every lane touches only its own column: no
// bank conflicts, no synchronisation), next to the operator T (nt^2 doubles, staged once per CTA).  Only the accumulators
// t = T r stay in registers.  One CTA of 4 warps per SM (shared memory bound): each warp feeds its own tensor pipe.
// Compared with the streamed kernel this removes the per-iteration launch, the HBM round trip of the state and the host
// synchronisation per check for exactly the horizon range where those dominate (config 4, H = 35 .. 60 at nu = 2).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "admm_onchip.cuh"

namespace mpcb {

// W warps per CTA, one CTA per SM: 8 (two warps per scheduler, so that one warp's elementwise phase hides under the other's DMMAs:
// 0.76 vs 0.57 of the pipe in the steady-state measurement of DESIGN 5.1) wherever T and 8 state slices fit 227 KB (NT <= 96 for
// sigma = 0), else 6, 5 or 4 (admm_smem.cu).

// shared memory: T fragments NT*NT, lo/hi and their integer keys (admm_onchip.cuh: dkey) NT each, per-warp parameter staging [8][npad], per-warp state (3 or 4) x KS x 32
__host__ __device__ inline size_t smemk_bytes(int NT, int np, bool sig, int W) {
  const int npad = (np + 1) & ~1;
  return sizeof(double) * ((size_t)NT * NT + 4 * NT + (size_t)W * 8 * npad + (size_t)W * (sig ? 4 : 3) * (NT / 4) * 32);
}

template <int NT, bool SIG, int W>
__global__ void __launch_bounds__(W * 32, 1) admm_smem_kernel(const OnchipParams P) {
  constexpr int SMEMK_WARPS = W, SMEMK_THREADS = W * 32;
  constexpr int EPL = NT / 4, KS = NT / 4, NTL = NT / 8;
  extern __shared__ __align__(16) double smem[];
  double* sT = smem;
  double* sLo = sT + NT * NT;
  double* sHi = sLo + NT;
  const int npad = (P.np + 1) & ~1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, l4 = lane & 3;
  long long* sLoK = reinterpret_cast<long long*>(sHi + NT);
  long long* sHiK = sLoK + NT;
  double* sP = sHi + 3 * NT + warp * 8 * npad;
  double* st = sHi + 3 * NT + SMEMK_WARPS * 8 * npad + (size_t)warp * (SIG ? 4 : 3) * KS * 32 + lane;   // this lane's column
  double* sC = st;                 // c = (1 - alpha) z + y / rho
  double* sR = st + KS * 32;       // r = rho (z - y/rho) + sigma x - q : the next MMA operand
  double* sQ = st + 2 * KS * 32;
  double* sX = st + 3 * KS * 32;   // only when SIG

  for (int i = threadIdx.x; i < NT * NT; i += SMEMK_THREADS) sT[i] = P.Tfrag[i];
  for (int i = threadIdx.x; i < NT; i += SMEMK_THREADS) { sLo[i] = P.lo[i]; sHi[i] = P.hi[i]; sLoK[i] = dkey(P.lo[i]); sHiK[i] = dkey(P.hi[i]); }
  __syncthreads();

  const double sigma = P.sigma, alpha = P.alpha, oma = 1.0 - P.alpha;
  const double rho_s = P.rho_box, rinv_s = 1.0 / P.rho_box, sig_rho = P.sigma + P.rho_box;
  const int nz = P.nz;
#pragma unroll 1
  for (int le = 0; le < EPL; le++) { sC[le * 32] = 0.0; sR[le * 32] = 0.0; sQ[le * 32] = 0.0; if (SIG) sX[le * 32] = 0.0; }
  long long pi = -1;
  int it_s = 0;
  double qn = 0.0;
  bool exhausted = false;
  const int max_iter = ((P.max_iter + P.check_every - 1) / P.check_every) * P.check_every;

  while (true) {
    // ------------------------------------------------------------------ refill finished / empty slots
    const bool need = (pi < 0) && !exhausted;
    if (__any_sync(0xffffffffu, need)) {
      long long np_i = -1;
      if (need && l4 == 0) np_i = (long long)atomicAdd(P.counter, 1ULL);
      np_i = __shfl_sync(0xffffffffu, np_i, lane & ~3);
      const bool fresh = need && np_i < P.batch;
      if (need && !fresh) exhausted = true;
      if (fresh) {
        pi = np_i;
        it_s = 0;
        for (int j = l4; j < P.np; j += 4) {
          double v;
          if (j < P.nx) v = P.x0[pi * P.nx + j];
          else if (j < 2 * P.nx) v = P.xref[(P.xref_bc ? 0 : pi) * P.nx + (j - P.nx)];
          else v = P.uref[(P.uref_bc ? 0 : pi) * P.nu + (j - 2 * P.nx)];
          sP[g * npad + j] = v;
        }
      }
      __syncwarp();
      if (fresh) {
        double m = 0.0;
#pragma unroll 1
        for (int t = 0; t < NTL; t++) {
          double a0 = 0.0, a1 = 0.0;       // q = Lq p for the two rows of this n-tile (Lt is read from global / L2: [np][NT])
          for (int j = 0; j < P.np; j++) {
            const double pj = sP[g * npad + j];
            const double2 l2 = *reinterpret_cast<const double2*>(&P.Lt[j * NT + 8 * t + 2 * l4]);
            a0 = fma(l2.x, pj, a0); a1 = fma(l2.y, pj, a1);
          }
          double u0 = 0.0, u1 = 0.0;       // settings.cold_init: v_unc = Lv p, the unconstrained optimum
          const bool cold_pt = P.Lv != nullptr && P.warm_v == nullptr;
          if (cold_pt) {
            for (int j = 0; j < P.np; j++) {
              const double pj = sP[g * npad + j];
              const double2 l2 = __ldg(reinterpret_cast<const double2*>(&P.Lv[j * NT + 8 * t + 2 * l4]));
              u0 = fma(l2.x, pj, u0); u1 = fma(l2.y, pj, u1);
            }
          }
#pragma unroll
          for (int jj = 0; jj < 2; jj++) {
            const int le = 2 * t + jj, e = 8 * t + 2 * l4 + jj;
            const double qv = jj ? a1 : a0;
            m = dmaxf(m, fabs(qv));
            double v0 = 0.0, ys0 = 0.0;    // OSQP warm start: x = v0, z = A x, y = y0
            if (P.warm_v != nullptr) {
              if (e < nz) v0 = P.warm_v[pi * nz + e];
              if (e < P.nt) ys0 = P.warm_y[pi * P.nt + e] * rinv_s;
            } else if (cold_pt && e < nz) {
              const double vu = jj ? u1 : u0;
              v0 = dclamp(vu, sLo[e], sHi[e]);
              ys0 = -MPCB_INIT_KAPPA * (v0 - vu);
            }
            sQ[le * 32] = qv;
            sC[le * 32] = fma(oma, v0, ys0);
            sR[le * 32] = fma(rho_s, v0 - ys0, fma(sigma, v0, -qv));
            if (SIG) sX[le * 32] = v0;
          }
        }
        qn = m;
      }
      qn = quad_max(qn);
    }
    if (!__any_sync(0xffffffffu, pi >= 0)) break;

    // ------------------------------------------------------------------ check_every ADMM iterations, the last one checks
    unsigned long long urp = 0ULL, urd = 0ULL, unA = 0ULL, unD = 0ULL;     // maxima of absolute values as bit patterns (integer pipe)
    double t[EPL];
    for (int ii = 0; ii < P.check_every; ii++) {
      const bool chk = (ii == P.check_every - 1);
#pragma unroll
      for (int i = 0; i < EPL; i++) t[i] = 0.0;
      const double* tp = sT + lane;
#pragma unroll 2
      for (int s = 0; s < KS; s++) {
        const double a = sR[s * 32];
#pragma unroll
        for (int tn = 0; tn < NTL; tn++) dmma884(t[2 * tn], t[2 * tn + 1], a, tp[(s * NTL + tn) * 32]);
      }
#pragma unroll
      for (int tn = 0; tn < NTL; tn++) {
        const double2 lo2 = *reinterpret_cast<const double2*>(&sLo[8 * tn + 2 * l4]);
        const double2 hi2 = *reinterpret_cast<const double2*>(&sHi[8 * tn + 2 * l4]);
        const longlong2 kl2 = *reinterpret_cast<const longlong2*>(&sLoK[8 * tn + 2 * l4]);
        const longlong2 kh2 = *reinterpret_cast<const longlong2*>(&sHiK[8 * tn + 2 * l4]);
#pragma unroll
        for (int jj = 0; jj < 2; jj++) {
          const int le = 2 * tn + jj;
          const double cv = sC[le * 32], qv = sQ[le * 32];
          const double w = fma(alpha, t[le], cv);
          const double zn = iclamp_k(w, jj ? lo2.y : lo2.x, jj ? hi2.y : hi2.x, jj ? kl2.y : kl2.x, jj ? kh2.y : kh2.x);
          if (chk) {   // residuals of (x~, z+, y+): Pc x~ = r - (sigma + rho) x~
            const double pc = fma(-sig_rho, t[le], sR[le * 32]);
            const double yb = rho_s * (w - zn);
            urp = umax64(urp, absbits(t[le] - zn));
            urd = umax64(urd, absbits(pc + qv + yb));
            unA = umax64(unA, umax64(absbits(t[le]), absbits(zn)));
            unD = umax64(unD, umax64(absbits(pc), absbits(yb)));
          }
          sC[le * 32] = fma(-alpha, zn, w);
          const double d = fma(2.0, zn, -w);
          if (SIG) {
            const double xn = fma(alpha, t[le], oma * sX[le * 32]);
            sX[le * 32] = xn;
            sR[le * 32] = fma(rho_s, d, fma(sigma, xn, -qv));
          } else {
            sR[le * 32] = fma(rho_s, d, -qv);
          }
        }
      }
    }
    it_s += P.check_every;

    // ------------------------------------------------------------------ termination (OSQP criteria at x~, z+, y+)
    double rp = __longlong_as_double((long long)urp), rd = __longlong_as_double((long long)urd), nA = __longlong_as_double((long long)unA), nD = __longlong_as_double((long long)unD);
    rp = quad_max(rp); rd = quad_max(rd); nA = quad_max(nA); nD = quad_max(nD);
    const bool conv = (rp <= P.eps_abs + P.eps_rel * nA) && (rd <= P.eps_abs + P.eps_rel * dmaxf(nD, qn));
    const bool fin = (pi >= 0) && (conv || it_s >= max_iter);
    if (fin) {
#pragma unroll
      for (int tn = 0; tn < NTL; tn++) {
        const int e = 8 * tn + 2 * l4;
        const double x0v = t[2 * tn], x1v = t[2 * tn + 1];          // x~ of the checking iteration is still in the accumulators
        if (((nz & 1) == 0) && e + 1 < nz) {
          *reinterpret_cast<double2*>(&P.v_out[pi * nz + e]) = make_double2(x0v, x1v);
        } else {
          if (e < nz) P.v_out[pi * nz + e] = x0v;
          if (e + 1 < nz) P.v_out[pi * nz + e + 1] = x1v;
        }
        if (P.y_out != nullptr) {
          // y+ from the updated state: c+ = w - alpha z+, r+ = rho (2 z+ - w) + sigma x+ - q  =>  z+ = (c+ + (r+ + q - sigma x+)/rho) / (2 - alpha)
#pragma unroll
          for (int jj = 0; jj < 2; jj++) {
            const int le = 2 * tn + jj;
            if (e + jj < P.nt) {
              const double cp = sC[le * 32];
              const double a2 = (sR[le * 32] + sQ[le * 32] - (SIG ? sigma * sX[le * 32] : 0.0)) * rinv_s;
              const double zn = (a2 + cp) / (2.0 - alpha);
              P.y_out[pi * P.nt + e + jj] = rho_s * (fma(alpha, zn, cp) - zn);
            }
          }
        }
      }
      if (l4 == 0) {
        P.status[pi] = conv ? 1 : -2;
        P.iters[pi] = it_s;
        P.pres[pi] = rp;
        P.dres[pi] = rd;
      }
      pi = -1;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long prev = atomicAdd(P.counter + 1, 1ULL);
    if (prev == (unsigned long long)gridDim.x - 1ULL) { P.counter[0] = 0ULL; P.counter[1] = 0ULL; }
  }
}

}  // namespace mpcb
