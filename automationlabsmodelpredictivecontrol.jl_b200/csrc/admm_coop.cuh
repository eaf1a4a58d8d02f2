// CTA-cooperative ADMM for FEW problems (sm_100a, round 2): the second rung of the rho ladder, and every small batch (mpcb_api.cu: up to 8 problems per SM,
// 32 where it replaces a shared-memory resident kernel).  First the general-row kernel, then its box-only counterpart.
//
// The slot kernels (admm_onchip.cuh, admm_smemg.cuh) give one warp eight problems and the whole operator product: right for throughput,
// wrong for the stragglers of a state-box batch -- a few dozen problems that need 1 000 .. 2 000 more iterations each, on a device that is
// otherwise idle.  Their time is one warp's iteration latency (nt = 120: 450 dependent-issue DMMAs + 30 rows of elementwise work ~ 8.8 us).
// Here ONE CTA of eight warps works on ONE group of eight problems: every warp owns an eighth of the n-tiles (output rows) of the product and
// of the per-row state, the operand vector r of the next iteration is exchanged through a double-buffered shared-memory slice, one CTA
// barrier per iteration.  Same arithmetic per row in the same order as the slot kernels (the k-step order of every accumulator is
// unchanged), so results are bit-identical to theirs; only the reductions of a check cross warps (maxima: order-free; the support sum of
// the infeasibility certificate is summed per warp, then over the warps).
// No ball rows (the ladder applies to state-box rows only); launched behind the first rung with the ticket count read from device memory,
// and it leaves the work to the slot kernel when that count is large (P.tickets_max).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "admm_onchip.cuh"

namespace mpcb {

constexpr int COOP_WARPS = 8;

// shared memory: T fragments NT*NT, lo / hi / rho / 1/rho NT each, parameter staging [8][npad], CTA-wide row slices (z, y/rho, q|b, [x], operand x2, check operand)
// of KS*32 doubles each, reduction scratch [warps][8 slots][8 values], one control word
__host__ __device__ inline size_t coop_bytes(int NT, int np, bool sig) {
  const int npad = (np + 1) & ~1;
  return sizeof(double) * ((size_t)NT * NT + 4 * NT + (size_t)8 * npad + (size_t)(sig ? 7 : 6) * (NT / 4) * 32 + COOP_WARPS * 8 * 8 + 2);
}

template <int NT, bool SIG>
__global__ void __launch_bounds__(COOP_WARPS * 32, 1) admm_coop_kernel(const OnchipParams P) {
  constexpr int W = COOP_WARPS, THREADS = W * 32;
  constexpr int KS = NT / 4, NTL = NT / 8, TPW = (NTL + W - 1) / W, RPW = 2 * TPW;   // tiles / rows per warp
  extern __shared__ __align__(16) double smem[];
  double* sT = smem;
  double* sLo = sT + NT * NT;
  double* sHi = sLo + NT;
  double* sRho = sHi + NT;
  double* sRinv = sRho + NT;
  const int npad = (P.np + 1) & ~1;
  double* sPar = sRinv + NT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, l4 = lane & 3;
  double* sl = sPar + 8 * npad + lane;          // this lane's column of the CTA-wide [row][lane] slices
  double* sZ = sl;
  double* sYs = sl + KS * 32;
  double* sQ = sl + 2 * KS * 32;
  double* sA0 = sl + 3 * KS * 32;               // operand r of the current / next iteration
  double* sA1 = sl + 4 * KS * 32;
  double* sIn = sl + 5 * KS * 32;               // operand of the check passes
  double* sX = sl + 6 * KS * 32;                // only when SIG
  double* sRed = sPar + 8 * npad + (size_t)(SIG ? 7 : 6) * KS * 32;      // [W][8][8]
  unsigned long long* sCtl = reinterpret_cast<unsigned long long*>(sRed + W * 8 * 8);

  for (int i = threadIdx.x; i < NT * NT; i += THREADS) sT[i] = P.Tfrag[i];
  for (int i = threadIdx.x; i < NT; i += THREADS) { sLo[i] = P.lo[i]; sHi[i] = P.hi[i]; sRho[i] = P.rho[i]; sRinv[i] = P.rinv[i]; }
  __syncthreads();

  const double sigma = P.sigma, alpha = P.alpha, oma = 1.0 - P.alpha;
  const int nz = P.nz;
  const int max_iter = ((P.max_iter + P.check_every - 1) / P.check_every) * P.check_every;
  long long batch_eff = P.batch_dev ? (long long)*P.batch_dev : P.batch;
  if (P.tickets_max >= 0 && batch_eff > P.tickets_max) batch_eff = 0;        // a large second rung stays with the slot kernel
  const int tn0 = warp * TPW;                   // first n-tile of this warp

  // out = (operator given by `frag`) * (the CTA-wide operand slice `op`), this warp's tiles only
  auto product = [&](const double* frag, const double* op, bool from_global, double (&out)[RPW]) {
#pragma unroll
    for (int i = 0; i < RPW; i++) out[i] = 0.0;
    const double* fp = frag + lane;
    if (from_global) {        // the check operator streams from L2: six k-steps of loads in flight per round trip
#pragma unroll 6
      for (int s = 0; s < KS; s++) {
        const double a = op[s * 32];
#pragma unroll
        for (int j = 0; j < TPW; j++)
          if (tn0 + j < NTL) dmma884(out[2 * j], out[2 * j + 1], a, __ldg(fp + (s * NTL + tn0 + j) * 32));
      }
    } else {
#pragma unroll 2
      for (int s = 0; s < KS; s++) {
        const double a = op[s * 32];
#pragma unroll
        for (int j = 0; j < TPW; j++)
          if (tn0 + j < NTL) dmma884(out[2 * j], out[2 * j + 1], a, fp[(s * NTL + tn0 + j) * 32]);
      }
    }
  };
  // per-slot combination of per-warp partials: every thread ends with the same value for its slot
  auto put = [&](int k, double v) { if (l4 == 0) sRed[(warp * 8 + g) * 8 + k] = v; };
  auto get_max = [&](int k) { double m = sRed[g * 8 + k]; for (int w = 1; w < W; w++) m = dmaxf(m, sRed[(w * 8 + g) * 8 + k]); return m; };
  auto get_sum = [&](int k) { double m = sRed[g * 8 + k]; for (int w = 1; w < W; w++) m += sRed[(w * 8 + g) * 8 + k]; return m; };

  for (;;) {
    // ------------------------------------------------------------------ next group of eight problems
    if (threadIdx.x == 0) sCtl[0] = atomicAdd(P.counter, 1ULL);
    __syncthreads();
    const long long base = (long long)sCtl[0] * 8;
    if (base >= batch_eff) break;
    long long pi = -1;
    if (base + g < batch_eff) pi = P.remap ? (long long)P.remap[base + g] : base + g;
    if (warp == 0 && pi >= 0) {
      for (int j = l4; j < P.np; j += 4) {
        double v;
        if (j < P.nx) v = P.x0[pi * P.nx + j];
        else if (j < 2 * P.nx) v = P.xref[(P.xref_bc ? 0 : pi) * P.nx + (j - P.nx)];
        else v = P.uref[(P.uref_bc ? 0 : pi) * P.nu + (j - 2 * P.nx)];
        sPar[g * npad + j] = v;
      }
    }
    __syncthreads();
    const bool cold_pt = P.Lv != nullptr && P.warm_v == nullptr;
    const bool start_pt = P.warm_v != nullptr || cold_pt;
    double m = 0.0;
#pragma unroll 1
    for (int j = 0; j < TPW; j++) {
      const int t = tn0 + j;
      if (t >= NTL) break;
      double a0 = 0.0, a1 = 0.0, u0 = 0.0, u1 = 0.0;
      if (pi >= 0) {
        for (int k = 0; k < P.np; k++) {
          const double pk = sPar[g * npad + k];
          const double2 l2 = *reinterpret_cast<const double2*>(&P.Lt[k * NT + 8 * t + 2 * l4]);
          a0 = fma(l2.x, pk, a0); a1 = fma(l2.y, pk, a1);
        }
        if (cold_pt)
          for (int k = 0; k < P.np; k++) {
            const double pk = sPar[g * npad + k];
            const double2 l2 = __ldg(reinterpret_cast<const double2*>(&P.Lv[k * NT + 8 * t + 2 * l4]));
            u0 = fma(l2.x, pk, u0); u1 = fma(l2.y, pk, u1);
          }
      }
#pragma unroll
      for (int jj = 0; jj < 2; jj++) {
        const int le = 2 * t + jj, e = 8 * t + 2 * l4 + jj;
        const bool box = e < nz;
        const double qv = jj ? a1 : a0;
        if (box) m = dmaxf(m, fabs(qv));
        double v0 = 0.0, ys0 = 0.0;
        if (pi >= 0) {
          if (P.warm_v != nullptr) {
            if (box) v0 = P.warm_v[pi * nz + e];
            if (e < P.nt) ys0 = P.warm_y[pi * P.nt + e] * sRinv[e];
          } else if (cold_pt && box) {
            const double vu = jj ? u1 : u0;
            v0 = dclamp(vu, sLo[e], sHi[e]);
            ys0 = -MPCB_INIT_KAPPA * (v0 - vu);
          }
        }
        sQ[le * 32] = qv;
        sZ[le * 32] = box ? v0 : 0.0;
        sYs[le * 32] = ys0;
        if (SIG) sX[le * 32] = box ? v0 : 0.0;
        sIn[le * 32] = box ? v0 : 0.0;
      }
    }
    put(0, quad_max(m));
    __syncthreads();
    const double qn = get_max(0);
    if (start_pt) {              // z_g = G x of the starting point: one pass with C
      double out[RPW];
      product(P.Cfrag, sIn, true, out);
#pragma unroll
      for (int j = 0; j < TPW; j++)
#pragma unroll
        for (int jj = 0; jj < 2; jj++) {
          const int t = tn0 + j, e = 8 * t + 2 * l4 + jj;
          if (t < NTL && e >= nz) sZ[(2 * t + jj) * 32] = out[2 * j + jj];
        }
    }
    // operand of the first iteration
#pragma unroll
    for (int j = 0; j < TPW; j++)
#pragma unroll
      for (int jj = 0; jj < 2; jj++) {
        const int t = tn0 + j, le = 2 * t + jj, e = 8 * t + 2 * l4 + jj;
        if (t < NTL) {
          double a = sRho[e] * (sZ[le * 32] - sYs[le * 32]);
          if (e < nz) a += SIG ? fma(sigma, sX[le * 32], -sQ[le * 32]) : -sQ[le * 32];
          sA0[le * 32] = a;
        }
      }
    __syncthreads();

    int it_s = 0;
    int cur = 0;
    for (;;) {
      // ---------------------------------------------------------------- check_every iterations, the last one checks
      double rp = 0.0, nA = 0.0, ndy = 0.0, supp = 0.0;
      double t[RPW], dys[RPW];
      for (int ii = 0; ii < P.check_every; ii++) {
        const bool chk = (ii == P.check_every - 1);
        const double* Ac = cur ? sA1 : sA0;
        double* An = cur ? sA0 : sA1;
        product(sT, Ac, false, t);
#pragma unroll
        for (int j = 0; j < TPW; j++) {
          const int tn = tn0 + j;
          if (tn < NTL) {
            const double2 lo2 = *reinterpret_cast<const double2*>(&sLo[8 * tn + 2 * l4]);
            const double2 hi2 = *reinterpret_cast<const double2*>(&sHi[8 * tn + 2 * l4]);
            const double2 rh2 = *reinterpret_cast<const double2*>(&sRho[8 * tn + 2 * l4]);
#pragma unroll
            for (int jj = 0; jj < 2; jj++) {
              const int le = 2 * tn + jj, e = 8 * tn + 2 * l4 + jj, li = 2 * j + jj;
              const bool box = e < nz;
              const double zp = sZ[le * 32], yp = sYs[le * 32], qv = sQ[le * 32];
              const double at = alpha * t[li];
              const double w = fma(oma, zp, at) + yp;
              double lo_e = jj ? lo2.y : lo2.x, hi_e = jj ? hi2.y : hi2.x;
              double xn = 0.0;
              if (box) { if (SIG) { xn = fma(oma, sX[le * 32], at); sX[le * 32] = xn; } }
              else { lo_e += qv; hi_e += qv; }
              const double zn = dclamp(w, lo_e, hi_e);
              const double yn = w - zn;
              const double rho_e = jj ? rh2.y : rh2.x;
              if (chk) {
                const double dy = rho_e * (yn - yp);
                dys[li] = dy;
                ndy = dmaxf(ndy, fabs(dy));
                supp += hi_e * dmaxf(dy, 0.0) + lo_e * (dy < 0.0 ? dy : 0.0);
                rp = dmaxf(rp, fabs(t[li] - zn));
                nA = dmaxf(nA, dmaxf(fabs(t[li]), fabs(zn)));
                sIn[le * 32] = box ? t[li] : rho_e * yn;           // [x~; y_g+]
              }
              sZ[le * 32] = zn;
              sYs[le * 32] = yn;
              double a = rho_e * (zn - yn);
              if (box) a += SIG ? fma(sigma, xn, -qv) : -qv;
              An[le * 32] = a;
            }
          }
        }
        __syncthreads();
        cur ^= 1;
      }
      it_s += P.check_every;

      // ---------------------------------------------------------------- termination (OSQP criteria at x~, z+, y+)
      // The dual residual needs a pass with C; a slot can only terminate on it when its primal residual has converged (or at the cap, or with an
      // infeasibility candidate, whose reported residuals must be complete): while no slot of the group is there, the pass is skipped.
      double rd = 0.0, nD = 0.0;
      put(0, quad_max(rp)); put(2, quad_max(nA)); put(4, quad_max(ndy)); put(5, quad_sum(supp));
      __syncthreads();
      rp = get_max(0); nA = get_max(2); ndy = get_max(4); supp = get_sum(5);
      const bool cand0 = (pi >= 0) && (ndy > P.eps_pinf) && (supp < -P.eps_pinf * ndy);
      const bool want_rd = (pi >= 0) && ((rp <= P.eps_abs + P.eps_rel * nA) || cand0 || it_s >= max_iter);
      if (__syncthreads_or(want_rd ? 1 : 0)) {
        double cc[RPW];
        product(P.Cfrag, sIn, true, cc);                           // [Pc x~ + G' y_g ; G x~], this warp's rows
#pragma unroll
        for (int j = 0; j < TPW; j++)
#pragma unroll
          for (int jj = 0; jj < 2; jj++) {
            const int tn = tn0 + j, le = 2 * tn + jj, e = 8 * tn + 2 * l4 + jj;
            if (tn < NTL && e < nz) {
              const double yb = sRho[e] * sYs[le * 32];
              rd = dmaxf(rd, fabs(cc[2 * j + jj] + sQ[le * 32] + yb));
              nD = dmaxf(nD, dmaxf(fabs(cc[2 * j + jj]), fabs(yb)));
            }
          }
        put(1, quad_max(rd)); put(3, quad_max(nD));
        __syncthreads();
        rd = get_max(1); nD = get_max(3);
      }
      const bool conv = want_rd && (rp <= P.eps_abs + P.eps_rel * nA) && (rd <= P.eps_abs + P.eps_rel * dmaxf(nD, qn));
      const bool cand = cand0 && !conv;
      bool pinf = false;
      if (__syncthreads_or(cand ? 1 : 0)) {        // (also the barrier between reading and rewriting the reduction scratch and sIn)
#pragma unroll
        for (int j = 0; j < TPW; j++)
#pragma unroll
          for (int jj = 0; jj < 2; jj++) {
            const int tn = tn0 + j, le = 2 * tn + jj, e = 8 * tn + 2 * l4 + jj;
            if (tn < NTL) sIn[le * 32] = (e < nz) ? 0.0 : dys[2 * j + jj];
          }
        __syncthreads();
        double cc[RPW];
        product(P.Cfrag, sIn, true, cc);
        double atdy = 0.0;
#pragma unroll
        for (int j = 0; j < TPW; j++)
#pragma unroll
          for (int jj = 0; jj < 2; jj++) {
            const int tn = tn0 + j, e = 8 * tn + 2 * l4 + jj;
            if (tn < NTL && e < nz) atdy = dmaxf(atdy, fabs(cc[2 * j + jj] + dys[2 * j + jj]));
          }
        put(6, quad_max(atdy));
        __syncthreads();
        atdy = get_max(6);
        pinf = cand && (atdy <= P.eps_pinf * ndy);
      }
      const bool fin = (pi >= 0) && (conv || pinf || it_s >= max_iter);
      if (fin) {
#pragma unroll
        for (int j = 0; j < TPW; j++) {
          const int tn = tn0 + j, e = 8 * tn + 2 * l4;
          if (tn < NTL) {
            if (e < nz) P.v_out[pi * nz + e] = t[2 * j];
            if (e + 1 < nz) P.v_out[pi * nz + e + 1] = t[2 * j + 1];
            if (P.y_out != nullptr) {
#pragma unroll
              for (int jj = 0; jj < 2; jj++)
                if (e + jj < P.nt) P.y_out[pi * P.nt + e + jj] = sRho[e + jj] * sYs[(2 * tn + jj) * 32];
            }
          }
        }
        if (warp == 0 && l4 == 0) {
          P.status[pi] = conv ? 1 : (pinf ? -3 : -2);
          P.iters[pi] = it_s + P.iters_add;
          P.pres[pi] = rp;
          P.dres[pi] = rd;
        }
        pi = -1;
      }
      if (__syncthreads_and(pi < 0 ? 1 : 0)) break;      // the whole group is done (also orders this check's scratch reads before the next writes)
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long prev = atomicAdd(P.counter + 1, 1ULL);
    if (prev == (unsigned long long)gridDim.x - 1ULL) { P.counter[0] = 0ULL; P.counter[1] = 0ULL; }
  }
}

// ---- box-only counterpart: small batches on controllers WITHOUT general rows (the reference's own closed-loop use: one plant, one problem per call).
// Per-row arithmetic of the box-only slot kernels (admm_onchip.cuh / admm_smem.cuh: c = (1 - alpha) z + y / rho and the next operand r per row,
// closed-form dual residual Pc x~ = r - (sigma + rho) x~, no pass with C), same k-step order per accumulator: results are bit-identical to theirs.
// shared memory: T fragments, lo / hi NT each, parameter staging [8][npad], CTA-wide slices c, q, [x], operand x2, reduction scratch, control word
__host__ __device__ inline size_t coopb_bytes(int NT, int np, bool sig) {
  const int npad = (np + 1) & ~1;
  return sizeof(double) * ((size_t)NT * NT + 2 * NT + (size_t)8 * npad + (size_t)(sig ? 5 : 4) * (NT / 4) * 32 + COOP_WARPS * 8 * 8 + 2);
}

template <int NT, bool SIG>
__global__ void __launch_bounds__(COOP_WARPS * 32, 1) admm_coopb_kernel(const OnchipParams P) {
  constexpr int W = COOP_WARPS, THREADS = W * 32;
  constexpr int KS = NT / 4, NTL = NT / 8, TPW = (NTL + W - 1) / W, RPW = 2 * TPW;
  extern __shared__ __align__(16) double smem[];
  double* sT = smem;
  double* sLo = sT + NT * NT;
  double* sHi = sLo + NT;
  const int npad = (P.np + 1) & ~1;
  double* sPar = sHi + NT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, l4 = lane & 3;
  double* sl = sPar + 8 * npad + lane;
  double* sC = sl;                              // c = (1 - alpha) z + y / rho
  double* sQ = sl + KS * 32;
  double* sA0 = sl + 2 * KS * 32;               // operand r of the current / next iteration
  double* sA1 = sl + 3 * KS * 32;
  double* sX = sl + 4 * KS * 32;                // only when SIG
  double* sRed = sPar + 8 * npad + (size_t)(SIG ? 5 : 4) * KS * 32;
  unsigned long long* sCtl = reinterpret_cast<unsigned long long*>(sRed + W * 8 * 8);

  for (int i = threadIdx.x; i < NT * NT; i += THREADS) sT[i] = P.Tfrag[i];
  for (int i = threadIdx.x; i < NT; i += THREADS) { sLo[i] = P.lo[i]; sHi[i] = P.hi[i]; }
  __syncthreads();

  const double sigma = P.sigma, alpha = P.alpha, oma = 1.0 - P.alpha;
  const double rho_s = P.rho_box, rinv_s = 1.0 / P.rho_box, sig_rho = P.sigma + P.rho_box;
  const int nz = P.nz;
  const int max_iter = ((P.max_iter + P.check_every - 1) / P.check_every) * P.check_every;
  const long long batch_eff = P.batch;
  const int tn0 = warp * TPW;
  auto put = [&](int k, double v) { if (l4 == 0) sRed[(warp * 8 + g) * 8 + k] = v; };
  auto get_max = [&](int k) { double m = sRed[g * 8 + k]; for (int w = 1; w < W; w++) m = dmaxf(m, sRed[(w * 8 + g) * 8 + k]); return m; };

  for (;;) {
    if (threadIdx.x == 0) sCtl[0] = atomicAdd(P.counter, 1ULL);
    __syncthreads();
    const long long base = (long long)sCtl[0] * 8;
    if (base >= batch_eff) break;
    long long pi = (base + g < batch_eff) ? base + g : -1;
    // the warm start sits next to the parameters in (possibly page-locked host) memory: both round trips are started together
    double wv[RPW], wy[RPW];
#pragma unroll
    for (int j = 0; j < TPW; j++)
#pragma unroll
      for (int jj = 0; jj < 2; jj++) {
        const int e = 8 * (tn0 + j) + 2 * l4 + jj;
        wv[2 * j + jj] = 0.0; wy[2 * j + jj] = 0.0;
        if (P.warm_v != nullptr && pi >= 0 && tn0 + j < NTL) {
          if (e < nz) wv[2 * j + jj] = P.warm_v[pi * nz + e];
          if (e < P.nt) wy[2 * j + jj] = P.warm_y[pi * P.nt + e];
        }
      }
    if (warp == 0 && pi >= 0) {
      for (int j = l4; j < P.np; j += 4) {
        double v;
        if (j < P.nx) v = P.x0[pi * P.nx + j];
        else if (j < 2 * P.nx) v = P.xref[(P.xref_bc ? 0 : pi) * P.nx + (j - P.nx)];
        else v = P.uref[(P.uref_bc ? 0 : pi) * P.nu + (j - 2 * P.nx)];
        sPar[g * npad + j] = v;
      }
    }
    __syncthreads();
    const bool cold_pt = P.Lv != nullptr && P.warm_v == nullptr;
    double m = 0.0;
#pragma unroll
    for (int j = 0; j < TPW; j++) {
      const int t = tn0 + j;
      if (t >= NTL) continue;
      double a0 = 0.0, a1 = 0.0, u0 = 0.0, u1 = 0.0;
      if (pi >= 0) {
        for (int k = 0; k < P.np; k++) {
          const double pk = sPar[g * npad + k];
          const double2 l2 = *reinterpret_cast<const double2*>(&P.Lt[k * NT + 8 * t + 2 * l4]);
          a0 = fma(l2.x, pk, a0); a1 = fma(l2.y, pk, a1);
        }
        if (cold_pt)
          for (int k = 0; k < P.np; k++) {
            const double pk = sPar[g * npad + k];
            const double2 l2 = __ldg(reinterpret_cast<const double2*>(&P.Lv[k * NT + 8 * t + 2 * l4]));
            u0 = fma(l2.x, pk, u0); u1 = fma(l2.y, pk, u1);
          }
      }
#pragma unroll
      for (int jj = 0; jj < 2; jj++) {
        const int le = 2 * t + jj, e = 8 * t + 2 * l4 + jj;
        const double qv = jj ? a1 : a0;
        m = dmaxf(m, fabs(qv));
        double v0 = 0.0, ys0 = 0.0;
        if (pi >= 0) {
          if (P.warm_v != nullptr) {
            v0 = wv[2 * j + jj];
            ys0 = wy[2 * j + jj] * rinv_s;
          } else if (cold_pt && e < nz) {
            const double vu = jj ? u1 : u0;
            v0 = dclamp(vu, sLo[e], sHi[e]);
            ys0 = -MPCB_INIT_KAPPA * (v0 - vu);
          }
        }
        sQ[le * 32] = qv;
        sC[le * 32] = fma(oma, v0, ys0);
        sA0[le * 32] = fma(rho_s, v0 - ys0, fma(sigma, v0, -qv));
        if (SIG) sX[le * 32] = v0;
      }
    }
    put(0, quad_max(m));
    __syncthreads();
    const double qn = get_max(0);

    int it_s = 0, cur = 0;
    for (;;) {
      double rp = 0.0, rd = 0.0, nA = 0.0, nD = 0.0;
      double t[RPW], yb[RPW];
      for (int ii = 0; ii < P.check_every; ii++) {
        const bool chk = (ii == P.check_every - 1);
        const double* Ac = cur ? sA1 : sA0;
        double* An = cur ? sA0 : sA1;
#pragma unroll
        for (int i = 0; i < RPW; i++) t[i] = 0.0;
        {
          const double* fp = sT + lane;
#pragma unroll 2
          for (int s = 0; s < KS; s++) {
            const double a = Ac[s * 32];
#pragma unroll
            for (int j = 0; j < TPW; j++)
              if (tn0 + j < NTL) dmma884(t[2 * j], t[2 * j + 1], a, fp[(s * NTL + tn0 + j) * 32]);
          }
        }
#pragma unroll
        for (int j = 0; j < TPW; j++) {
          const int tn = tn0 + j;
          if (tn < NTL) {
            const double2 lo2 = *reinterpret_cast<const double2*>(&sLo[8 * tn + 2 * l4]);
            const double2 hi2 = *reinterpret_cast<const double2*>(&sHi[8 * tn + 2 * l4]);
#pragma unroll
            for (int jj = 0; jj < 2; jj++) {
              const int le = 2 * tn + jj, li = 2 * j + jj;
              const double qv = sQ[le * 32];
              const double w = fma(alpha, t[li], sC[le * 32]);
              const double zn = dclamp(w, jj ? lo2.y : lo2.x, jj ? hi2.y : hi2.x);
              if (chk) {   // residuals of (x~, z+, y+): Pc x~ = r - (sigma + rho) x~
                const double pc = fma(-sig_rho, t[li], Ac[le * 32]);
                const double y1 = rho_s * (w - zn);
                yb[li] = y1;
                rp = dmaxf(rp, fabs(t[li] - zn));
                rd = dmaxf(rd, fabs(pc + qv + y1));
                nA = dmaxf(nA, dmaxf(fabs(t[li]), fabs(zn)));
                nD = dmaxf(nD, dmaxf(fabs(pc), fabs(y1)));
              }
              sC[le * 32] = fma(-alpha, zn, w);
              const double d = fma(2.0, zn, -w);
              if (SIG) {
                const double xn = fma(alpha, t[li], oma * sX[le * 32]);
                sX[le * 32] = xn;
                An[le * 32] = fma(rho_s, d, fma(sigma, xn, -qv));
              } else {
                An[le * 32] = fma(rho_s, d, -qv);
              }
            }
          }
        }
        __syncthreads();
        cur ^= 1;
      }
      it_s += P.check_every;
      put(0, quad_max(rp)); put(1, quad_max(rd)); put(2, quad_max(nA)); put(3, quad_max(nD));
      __syncthreads();
      rp = get_max(0); rd = get_max(1); nA = get_max(2); nD = get_max(3);
      const bool conv = (rp <= P.eps_abs + P.eps_rel * nA) && (rd <= P.eps_abs + P.eps_rel * dmaxf(nD, qn));
      const bool fin = (pi >= 0) && (conv || it_s >= max_iter);
      if (fin) {
#pragma unroll
        for (int j = 0; j < TPW; j++) {
          const int tn = tn0 + j, e = 8 * tn + 2 * l4;
          if (tn < NTL) {
            if (e < nz) P.v_out[pi * nz + e] = t[2 * j];
            if (e + 1 < nz) P.v_out[pi * nz + e + 1] = t[2 * j + 1];
            if (P.y_out != nullptr) {
              if (e < P.nt) P.y_out[pi * P.nt + e] = yb[2 * j];
              if (e + 1 < P.nt) P.y_out[pi * P.nt + e + 1] = yb[2 * j + 1];
            }
          }
        }
        if (warp == 0 && l4 == 0) {
          P.status[pi] = conv ? 1 : -2;
          P.iters[pi] = it_s + P.iters_add;
          P.pres[pi] = rp;
          P.dres[pi] = rd;
        }
        pi = -1;
      }
      if (__syncthreads_and(pi < 0 ? 1 : 0)) break;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long prev = atomicAdd(P.counter + 1, 1ULL);
    if (prev == (unsigned long long)gridDim.x - 1ULL) { P.counter[0] = 0ULL; P.counter[1] = 0ULL; }
  }
}

}  // namespace mpcb
