// Shared-memory resident batched ADMM WITH general rows (terminal equality / contractive ball / state box) for 48 < nt <= 120 (compiled from
// nt8 = 32; the automatic choice from nt8 = 56, mpcb_api.cu) (sm_100a, round 2): the general-row counterpart of admm_smem.cuh.  Before it, every controller with general rows and nt > 64 went
// to the streamed GEMM path, whose check period costs ~0.2-0.5 ms of launches plus a host synchronisation however few rows are left --
// exactly wrong for the long iteration tails of active state-box rows (H = 20 with the state box: nt = 120, a few problems of 10^4 need
// thousands of iterations).  Here a problem never leaves the SM and the call never synchronises with the host.
//
// Same algorithm and slot scheme as the general-row instantiation of admm_onchip.cuh (one warp = 8 problem slots along the M dimension
// of DMMA.8x8x4, T in fragment order as the B operand, finished slots refilled from the global ticket counter at every check, the rho
// ladder's index map and device-side ticket count), with the data placement of admm_smem.cuh: the operator T (nt^2 doubles) is staged
// once per CTA, the per-row state z, y/rho, q|b (and x when sigma > 0) lives in a per-warp [row][lane] slice, only the accumulators stay
// in registers.  The check operator C = [[Pc, G'],[G, 0]] does not fit next to T (2 x 115 KB at nt = 120): the one or two passes a check
// makes with it read its fragments straight from global memory (L2-resident, identical for every warp of the grid); their A operands are
// indexed by the k-step of a rolled loop and therefore sit in per-thread local memory (L1), not in another shared-memory slice.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "admm_onchip.cuh"

namespace mpcb {

// shared memory: T fragments NT*NT, lo / hi / rho / 1/rho NT each, per-warp parameter staging [8][npad], per-warp state (3 or 4) x KS x 32
__host__ __device__ constexpr size_t smemg_bytes(int NT, int np, bool sig, int W) {
  const int npad = (np + 1) & ~1;
  return sizeof(double) * ((size_t)NT * NT + 4 * NT + (size_t)W * 8 * npad + (size_t)W * (sig ? 4 : 3) * (NT / 4) * 32);
}

template <int NT, bool SIG, int W>
__global__ void __launch_bounds__(W * 32, 1) admm_smemg_kernel(const OnchipParams P) {
  constexpr int THREADS = W * 32;
  constexpr int EPL = NT / 4, KS = NT / 4, NTL = NT / 8;
  extern __shared__ __align__(16) double smem[];
  double* sT = smem;
  double* sLo = sT + NT * NT;
  double* sHi = sLo + NT;
  double* sRho = sHi + NT;
  double* sRinv = sRho + NT;
  const int npad = (P.np + 1) & ~1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, l4 = lane & 3;
  double* sP = sRinv + NT + warp * 8 * npad;
  double* st = sRinv + NT + W * 8 * npad + (size_t)warp * (SIG ? 4 : 3) * KS * 32 + lane;   // this lane's column
  double* sZ = st;                  // z
  double* sYs = st + KS * 32;       // y / rho
  double* sQ = st + 2 * KS * 32;    // q (box rows) / bound offset b(p) (general rows)
  double* sX = st + 3 * KS * 32;    // relaxed x of the box rows, only when SIG

  for (int i = threadIdx.x; i < NT * NT; i += THREADS) sT[i] = P.Tfrag[i];
  for (int i = threadIdx.x; i < NT; i += THREADS) { sLo[i] = P.lo[i]; sHi[i] = P.hi[i]; sRho[i] = P.rho[i]; sRinv[i] = P.rinv[i]; }
  __syncthreads();

  const double sigma = P.sigma, alpha = P.alpha, oma = 1.0 - P.alpha;
  const int nz = P.nz;
  unsigned boxbits = 0;
#pragma unroll
  for (int le = 0; le < EPL; le++) {
    const int e = 8 * (le >> 1) + 2 * l4 + (le & 1);
    if (e < nz) boxbits |= 1u << le;
  }
#pragma unroll 1
  for (int le = 0; le < EPL; le++) { sZ[le * 32] = 0.0; sYs[le * 32] = 0.0; sQ[le * 32] = 0.0; if (SIG) sX[le * 32] = 0.0; }
  long long pi = -1;
  int it_s = 0;
  double qn = 0.0, rad = 0.0;
  bool exhausted = false;
  const int max_iter = ((P.max_iter + P.check_every - 1) / P.check_every) * P.check_every;
  long long batch_eff = P.batch_dev ? (long long)*P.batch_dev : P.batch;
  if (P.batch_dev && batch_eff <= P.tickets_skip_le) batch_eff = 0;

  // per-thread operands of the check passes, indexed by the k-step of a rolled loop: local memory
  double in_l[EPL], dys_l[EPL];

  // out = C in  with the fragments of C read from global memory (checks and warm starts only)
  auto c_pass = [&](double (&out)[EPL]) {
#pragma unroll
    for (int i = 0; i < EPL; i++) out[i] = 0.0;
    const double* cp = P.Cfrag + lane;
#pragma unroll 2
    for (int s = 0; s < KS; s++) {
      const double a = in_l[s];
#pragma unroll
      for (int tn = 0; tn < NTL; tn++) dmma884(out[2 * tn], out[2 * tn + 1], a, __ldg(cp + (s * NTL + tn) * 32));
    }
  };

  while (true) {
    // ------------------------------------------------------------------ refill finished / empty slots
    const bool need = (pi < 0) && !exhausted;
    if (__any_sync(0xffffffffu, need)) {
      long long np_i = -1;
      if (need && l4 == 0) np_i = (long long)atomicAdd(P.counter, 1ULL);
      np_i = __shfl_sync(0xffffffffu, np_i, lane & ~3);
      const bool fresh = need && np_i < batch_eff;
      if (need && !fresh) exhausted = true;
      if (fresh) {
        pi = P.remap ? (long long)P.remap[np_i] : np_i;
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
      if (P.nball > 0) {
        double d2 = 0.0;
        if (fresh)
          for (int j = l4; j < P.nx; j += 4) { const double dv = sP[g * npad + j] - sP[g * npad + P.nx + j]; d2 = fma(dv, dv, d2); }
        d2 = quad_sum(d2);
        if (fresh) rad = sqrt(0.9 * d2);
      }
      const bool cold_pt = P.Lv != nullptr && P.warm_v == nullptr;
      const bool start_pt = P.warm_v != nullptr || cold_pt;       // a starting point other than zero: z_g = G x needs one C pass
      if (fresh) {
        double m = 0.0;
#pragma unroll 1
        for (int t = 0; t < NTL; t++) {
          double a0 = 0.0, a1 = 0.0;       // q = Lq p (box rows) / b = Lb p (general rows) for the two rows of this n-tile (Lt read from global / L2)
          for (int j = 0; j < P.np; j++) {
            const double pj = sP[g * npad + j];
            const double2 l2 = *reinterpret_cast<const double2*>(&P.Lt[j * NT + 8 * t + 2 * l4]);
            a0 = fma(l2.x, pj, a0); a1 = fma(l2.y, pj, a1);
          }
          double u0 = 0.0, u1 = 0.0;       // settings.cold_init: v_unc = Lv p
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
            const bool box = e < nz;
            const double qv = jj ? a1 : a0;
            if (box) m = dmaxf(m, fabs(qv));
            double v0 = 0.0, ys0 = 0.0;    // OSQP warm start: x = v0, z = A x, y = y0
            if (P.warm_v != nullptr) {
              if (box) v0 = P.warm_v[pi * nz + e];
              if (e < P.nt) ys0 = P.warm_y[pi * P.nt + e] * sRinv[e];
            } else if (cold_pt && box) {
              const double vu = jj ? u1 : u0;
              v0 = dclamp(vu, sLo[e], sHi[e]);
              ys0 = -MPCB_INIT_KAPPA * (v0 - vu);
            }
            sQ[le * 32] = qv;
            sZ[le * 32] = box ? v0 : 0.0;
            sYs[le * 32] = ys0;
            if (SIG) sX[le * 32] = box ? v0 : 0.0;
          }
        }
        qn = m;
      }
      qn = quad_max(qn);
      if (start_pt) {
        double out[EPL];
#pragma unroll 1
        for (int le = 0; le < EPL; le++) in_l[le] = (fresh && ((boxbits >> le) & 1u)) ? sZ[le * 32] : 0.0;
        c_pass(out);
#pragma unroll
        for (int le = 0; le < EPL; le++)
          if (fresh && !((boxbits >> le) & 1u)) sZ[le * 32] = out[le];
      }
    }
    if (!__any_sync(0xffffffffu, pi >= 0)) break;

    // ------------------------------------------------------------------ check_every ADMM iterations, the last one checks
    double rp = 0.0, rd = 0.0, nA = 0.0, nD = 0.0, ndy = 0.0, supp = 0.0;
    double t[EPL];
    for (int ii = 0; ii < P.check_every; ii++) {
      const bool chk = (ii == P.check_every - 1);
#pragma unroll
      for (int i = 0; i < EPL; i++) t[i] = 0.0;
      const double* tp = sT + lane;
#pragma unroll 2
      for (int s = 0; s < KS; s++) {
        const int e = 8 * (s >> 1) + 2 * l4 + (s & 1);
        double a = sRho[e] * (sZ[s * 32] - sYs[s * 32]);         // r = rho (z - y/rho) [+ sigma x - q on the box rows]
        if (e < nz) a += SIG ? fma(sigma, sX[s * 32], -sQ[s * 32]) : -sQ[s * 32];
#pragma unroll
        for (int tn = 0; tn < NTL; tn++) dmma884(t[2 * tn], t[2 * tn + 1], a, tp[(s * NTL + tn) * 32]);
      }
      // contractive terminal set: the rows nz .. nz + nball - 1 of a slot are projected onto ONE ball (admm_onchip.cuh)
      double bscale = 1.0;
      if (P.nball > 0) {
        double d2 = 0.0;
#pragma unroll
        for (int le = 0; le < EPL; le++) {
          const int e = 8 * (le >> 1) + 2 * l4 + (le & 1);
          if (e >= nz && e < nz + P.nball) { const double dv = fma(oma, sZ[le * 32], alpha * t[le]) + sYs[le * 32] - sQ[le * 32]; d2 = fma(dv, dv, d2); }
        }
        d2 = quad_sum(d2);
        if (d2 > rad * rad) bscale = rad / sqrt(d2);
      }
#pragma unroll
      for (int tn = 0; tn < NTL; tn++) {
        const double2 lo2 = *reinterpret_cast<const double2*>(&sLo[8 * tn + 2 * l4]);
        const double2 hi2 = *reinterpret_cast<const double2*>(&sHi[8 * tn + 2 * l4]);
        const double2 rh2 = *reinterpret_cast<const double2*>(&sRho[8 * tn + 2 * l4]);
#pragma unroll
        for (int jj = 0; jj < 2; jj++) {
          const int le = 2 * tn + jj, e = 8 * tn + 2 * l4 + jj;
          const bool box = e < nz;
          const double zp = sZ[le * 32], yp = sYs[le * 32], qv = sQ[le * 32];
          const double at = alpha * t[le];
          const double w = fma(oma, zp, at) + yp;
          double lo_e = jj ? lo2.y : lo2.x, hi_e = jj ? hi2.y : hi2.x;
          if (box) { if (SIG) sX[le * 32] = fma(oma, sX[le * 32], at); }
          else { lo_e += qv; hi_e += qv; }
          double zn = dclamp(w, lo_e, hi_e);
          if (P.nball > 0 && e >= nz && e < nz + P.nball) zn = fma(w - qv, bscale, qv);
          const double yn = w - zn;
          if (chk) {
            const double rho_e = jj ? rh2.y : rh2.x;
            const double dy = rho_e * (yn - yp);
            dys_l[le] = dy;
            ndy = dmaxf(ndy, fabs(dy));
            supp += (hi_e) * dmaxf(dy, 0.0) + (lo_e) * (dy < 0.0 ? dy : 0.0);
            rp = dmaxf(rp, fabs(t[le] - zn));                    // A x~ = [x~; G x~] = t
            nA = dmaxf(nA, dmaxf(fabs(t[le]), fabs(zn)));
            in_l[le] = box ? t[le] : rho_e * yn;                 // [x~; y_g+]: operand of the residual pass
          }
          sZ[le * 32] = zn;
          sYs[le * 32] = yn;
        }
      }
    }
    it_s += P.check_every;

    // ------------------------------------------------------------------ termination (OSQP criteria at x~, z+, y+)
    // The dual residual needs a pass with C; a slot can only terminate on it when its primal residual has converged (or at the cap, or with an
    // infeasibility candidate, whose reported residuals must be complete): while no slot of the warp is there, the pass is skipped.
    bool pinf = false;
    rp = quad_max(rp); nA = quad_max(nA); ndy = quad_max(ndy); supp = quad_sum(supp);
    const bool prim_ok = rp <= P.eps_abs + P.eps_rel * nA;
    const bool cand0 = (pi >= 0) && (P.nball == 0) && (ndy > P.eps_pinf) && (supp < -P.eps_pinf * ndy);   // no certificate is evaluated for ball rows
    const bool want_rd = (pi >= 0) && (prim_ok || cand0 || it_s >= max_iter);
    if (__any_sync(0xffffffffu, want_rd)) {
      double cc[EPL];
      c_pass(cc);                                                 // [Pc x~ + G' y_g ; G x~]
#pragma unroll
      for (int tn = 0; tn < NTL; tn++) {
#pragma unroll
        for (int jj = 0; jj < 2; jj++) {
          const int le = 2 * tn + jj, e = 8 * tn + 2 * l4 + jj;
          if (e < nz) {
            const double yb = sRho[e] * sYs[le * 32];
            rd = dmaxf(rd, fabs(cc[le] + sQ[le * 32] + yb));
            nD = dmaxf(nD, dmaxf(fabs(cc[le]), fabs(yb)));
          }
        }
      }
    }
    rd = quad_max(rd); nD = quad_max(nD);
    const bool conv = want_rd && prim_ok && (rd <= P.eps_abs + P.eps_rel * dmaxf(nD, qn));
    {  // OSQP primal infeasibility certificate on delta_y of the last iteration
      const bool cand = cand0 && !conv;
      if (__any_sync(0xffffffffu, cand)) {
        double cc[EPL];
#pragma unroll 1
        for (int le = 0; le < EPL; le++) in_l[le] = ((boxbits >> le) & 1u) ? 0.0 : dys_l[le];
        c_pass(cc);
        double atdy = 0.0;
#pragma unroll
        for (int le = 0; le < EPL; le++)
          if ((boxbits >> le) & 1u) atdy = dmaxf(atdy, fabs(cc[le] + dys_l[le]));
        atdy = quad_max(atdy);
        pinf = cand && (atdy <= P.eps_pinf * ndy);
      }
    }
    const bool fin = (pi >= 0) && (conv || pinf || it_s >= max_iter);
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
#pragma unroll
          for (int jj = 0; jj < 2; jj++)
            if (e + jj < P.nt) P.y_out[pi * P.nt + e + jj] = sRho[e + jj] * sYs[(2 * tn + jj) * 32];
        }
      }
      if (l4 == 0) {
        P.status[pi] = conv ? 1 : (pinf ? -3 : -2);
        P.iters[pi] = it_s + P.iters_add;
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
