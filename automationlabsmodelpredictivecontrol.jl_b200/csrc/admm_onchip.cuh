// On-chip batched OSQP-style ADMM for the condensed linear-MPC QP (sm_100a).
//
// Replaces the per-problem `JuMP.optimize!` -> OSQP call (/root/reference/src/main/computation_mpc.jl:41) for
// thousands of problems that share one designed controller.  One warp owns 8 problem SLOTS (one per lane quad).
// Per ADMM iteration the multi-RHS contraction  [x~; z~_g] = T r  (T = [I;G] K^-1 [I,G'], cached per system) runs on
// the FP64 tensor pipe as DMMA.8x8x4 with problems along M, the fragment-ordered T streamed from shared memory as
// the B operand and the per-problem right-hand side produced in registers as the A operand.  Because T is symmetric
// and we control the row order of the B fragments, the C fragment of one iteration IS the A fragment layout of the
// next: the ADMM state (x, q, z, y/rho: 4 doubles per row) never leaves the register file between iterations, and
// projection / dual update / residual reductions are fused right behind the MMAs (quad shuffles only).
// Slots that terminate are written out and refilled from a global work counter at every check point, so iteration
// count variance across the batch (10x on the quadruple-tank batch) does not idle the tensor pipe.
//
// Row -> lane mapping: lane = 4*g + l4 (g = slot 0..7, l4 = 0..3); local row le = 2*t + j  <->  row e = 8*t + 2*l4 + j.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mpcb {

struct OnchipParams {
  // per-system constants (device memory)
  const double* Tfrag;  // NT*NT fragment order  [(s*NTL + tn)*32 + lane] = T[8*(s>>1) + 2*l4 + (s&1)][8*tn + g]
  const double* Cfrag;  // same order for C = [[Pc,G'],[G,0]]
  const double* Lt;     // [np][NT]  transposed parameter map  (rows < nz: Lq ; rows >= nz: Lb)
  const double* lo;     // [NT] constant part of lower bounds
  const double* hi;     // [NT]
  const double* rho;    // [NT] (padding rows: 1)
  const double* rinv;   // [NT]
  int nz, nt, np, nx, nu;
  double rho_box, sigma, alpha, eps_abs, eps_rel, eps_pinf;
  int max_iter, check_every;
  // batch
  long long batch;
  const double* x0;     // [batch][nx]
  const double* xref;   // [batch or 1][nx]
  const double* uref;   // [batch or 1][nu]
  int xref_bc, uref_bc;
  const double* warm_v; // [batch][nz] or null
  const double* warm_y; // [batch][nt] or null
  double* v_out;        // [batch][nz]  absolute inputs (OSQP's x)
  double* y_out;        // [batch][nt] or null
  int32_t* status;
  int32_t* iters;
  double* pres;
  double* dres;
  unsigned long long* counter;  // work queue head (zeroed before launch)
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ double quad_max(double v) {
  v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}
__device__ __forceinline__ double quad_sum(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

constexpr int ONCHIP_THREADS = 128;
constexpr int ONCHIP_WARPS = ONCHIP_THREADS / 32;

// shared memory: T, C fragments, Lt, lo, hi, rho, rinv, then per-warp parameter staging [8][npad]
__host__ __device__ inline size_t onchip_smem_bytes(int NT, int np) {
  int npad = (np + 1) & ~1;
  return sizeof(double) * ((size_t)2 * NT * NT + (size_t)np * NT + 4 * NT + (size_t)ONCHIP_WARPS * 8 * npad);
}

template <int NT, bool HAS_G, int MINB>
__global__ void __launch_bounds__(ONCHIP_THREADS, MINB) admm_onchip_kernel(const OnchipParams P) {
  constexpr int EPL = NT / 4;  // rows per lane
  constexpr int KS = NT / 4;   // k-steps
  constexpr int NTL = NT / 8;  // n-tiles
  extern __shared__ __align__(16) double smem[];
  double* sT = smem;
  double* sC = sT + NT * NT;
  double* sL = sC + NT * NT;
  double* sLo = sL + P.np * NT;
  double* sHi = sLo + NT;
  double* sRho = sHi + NT;
  double* sRinv = sRho + NT;
  const int npad = (P.np + 1) & ~1;
  double* sP = sRinv + NT + (threadIdx.x >> 5) * 8 * npad;

  for (int i = threadIdx.x; i < NT * NT; i += ONCHIP_THREADS) { sT[i] = P.Tfrag[i]; sC[i] = P.Cfrag[i]; }
  for (int i = threadIdx.x; i < P.np * NT; i += ONCHIP_THREADS) sL[i] = P.Lt[i];
  for (int i = threadIdx.x; i < NT; i += ONCHIP_THREADS) { sLo[i] = P.lo[i]; sHi[i] = P.hi[i]; sRho[i] = P.rho[i]; sRinv[i] = P.rinv[i]; }
  __syncthreads();

  const int lane = threadIdx.x & 31, g = lane >> 2, l4 = lane & 3;
  const double sigma = P.sigma, alpha = P.alpha, oma = 1.0 - P.alpha;
  const double rho_s = P.rho_box, rinv_s = 1.0 / P.rho_box;
  const int nz = P.nz;

  // ADMM state, one entry per owned row.  Box rows: x, q, z, ys (= y/rho).  General rows (HAS_G): the x slot holds
  // the per-problem bound offset b(p) and q is 0.
  double x[EPL], q[EPL], z[EPL], ys[EPL];
  double dys[HAS_G ? EPL : 1];
  unsigned boxbits = 0;
#pragma unroll
  for (int le = 0; le < EPL; le++) {
    const int e = 8 * (le >> 1) + 2 * l4 + (le & 1);
    if (!HAS_G || e < nz) boxbits |= 1u << le;
    x[le] = q[le] = z[le] = ys[le] = 0.0;
  }
  (void)dys;
  long long pi = -1;   // problem held by this slot (quad-uniform)
  int it_s = 0;        // its iteration count
  double qn = 0.0;     // ||q||_inf
  bool exhausted = false;
  const int max_iter = ((P.max_iter + P.check_every - 1) / P.check_every) * P.check_every;

  // one product  out = M in  with M in fragment order (sT or sC)
  auto mma_pass = [&](const double* __restrict__ sM, const double (&in)[EPL], double (&out)[EPL]) {
#pragma unroll
    for (int i = 0; i < EPL; i++) out[i] = 0.0;
#pragma unroll
    for (int s = 0; s < KS; s++) {
#pragma unroll
      for (int tn = 0; tn < NTL; tn++) dmma884(out[2 * tn], out[2 * tn + 1], in[s], sM[(s * NTL + tn) * 32 + lane]);
    }
  };

  while (true) {
    // ------------------------------------------------------------------ refill finished / empty slots
    bool need = (pi < 0) && !exhausted;
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
        double acc[EPL];
#pragma unroll
        for (int le = 0; le < EPL; le++) acc[le] = 0.0;
        for (int j = 0; j < P.np; j++) {
          const double pj = sP[g * npad + j];
#pragma unroll
          for (int t = 0; t < NTL; t++) {
            const double2 l2 = *reinterpret_cast<const double2*>(&sL[j * NT + 8 * t + 2 * l4]);
            acc[2 * t] = fma(l2.x, pj, acc[2 * t]);
            acc[2 * t + 1] = fma(l2.y, pj, acc[2 * t + 1]);
          }
        }
        double m = 0.0;
#pragma unroll
        for (int le = 0; le < EPL; le++) {
          const int e = 8 * (le >> 1) + 2 * l4 + (le & 1);
          const bool box = (boxbits >> le) & 1u;
          if (box) { q[le] = acc[le]; x[le] = 0.0; m = fmax(m, fabs(acc[le])); }
          else { q[le] = 0.0; x[le] = acc[le]; }
          z[le] = 0.0; ys[le] = 0.0;
          if (P.warm_v != nullptr) {
            if (box && e < nz) { x[le] = P.warm_v[pi * nz + e]; z[le] = x[le]; }
            if (e < P.nt) ys[le] = P.warm_y[pi * P.nt + e] * (HAS_G ? sRinv[e] : rinv_s);
          }
        }
        qn = m;
      }
      qn = quad_max(qn);
      if (HAS_G && P.warm_v != nullptr) {  // OSQP warm start sets z = A x: general rows need G x
        double in[EPL], out[EPL];
#pragma unroll
        for (int le = 0; le < EPL; le++) in[le] = (fresh && ((boxbits >> le) & 1u)) ? x[le] : 0.0;
        mma_pass(sC, in, out);
#pragma unroll
        for (int le = 0; le < EPL; le++)
          if (fresh && !((boxbits >> le) & 1u)) z[le] = out[le];
      }
    }
    if (!__any_sync(0xffffffffu, pi >= 0)) break;

    // ------------------------------------------------------------------ check_every ADMM iterations
    for (int ii = 0; ii < P.check_every; ii++) {
      const bool last = HAS_G && (ii == P.check_every - 1);
      double t[EPL];
#pragma unroll
      for (int i = 0; i < EPL; i++) t[i] = 0.0;
#pragma unroll
      for (int s = 0; s < KS; s++) {
        double r;
        if (!HAS_G) {
          r = fma(rho_s, z[s] - ys[s], fma(sigma, x[s], -q[s]));
        } else {
          const int e = 8 * (s >> 1) + 2 * l4 + (s & 1);
          r = sRho[e] * (z[s] - ys[s]);
          if ((boxbits >> s) & 1u) r += fma(sigma, x[s], -q[s]);
        }
#pragma unroll
        for (int tn = 0; tn < NTL; tn++) dmma884(t[2 * tn], t[2 * tn + 1], r, sT[(s * NTL + tn) * 32 + lane]);
      }
#pragma unroll
      for (int tn = 0; tn < NTL; tn++) {
        const double2 lo2 = *reinterpret_cast<const double2*>(&sLo[8 * tn + 2 * l4]);
        const double2 hi2 = *reinterpret_cast<const double2*>(&sHi[8 * tn + 2 * l4]);
#pragma unroll
        for (int jj = 0; jj < 2; jj++) {
          const int le = 2 * tn + jj;
          const double at = alpha * t[le];
          const double w = fma(oma, z[le], at) + ys[le];
          double lo_e = jj ? lo2.y : lo2.x, hi_e = jj ? hi2.y : hi2.x;
          if (HAS_G) {
            const bool box = (boxbits >> le) & 1u;
            if (box) x[le] = fma(oma, x[le], at);
            else { lo_e += x[le]; hi_e += x[le]; }
          } else {
            x[le] = fma(oma, x[le], at);
          }
          const double zn = fmin(fmax(w, lo_e), hi_e);
          const double yn = w - zn;
          if (HAS_G) { if (last) dys[le] = yn - ys[le]; }
          z[le] = zn;
          ys[le] = yn;
        }
      }
    }
    it_s += P.check_every;

    // ------------------------------------------------------------------ termination check (OSQP criteria)
    double rp = 0.0, rd = 0.0, nA = 0.0, nD = 0.0;
    {
      double in[EPL], c[EPL];
#pragma unroll
      for (int le = 0; le < EPL; le++) {
        if (!HAS_G) in[le] = x[le];
        else {
          const int e = 8 * (le >> 1) + 2 * l4 + (le & 1);
          in[le] = ((boxbits >> le) & 1u) ? x[le] : sRho[e] * ys[le];
        }
      }
      mma_pass(sC, in, c);
#pragma unroll
      for (int le = 0; le < EPL; le++) {
        const bool box = (boxbits >> le) & 1u;
        if (box) {
          const int e = 8 * (le >> 1) + 2 * l4 + (le & 1);
          const double y = (HAS_G ? sRho[e] : rho_s) * ys[le];
          rp = fmax(rp, fabs(x[le] - z[le]));
          rd = fmax(rd, fabs(c[le] + q[le] + y));
          nA = fmax(nA, fmax(fabs(x[le]), fabs(z[le])));
          nD = fmax(nD, fmax(fabs(c[le]), fabs(y)));
        } else {
          rp = fmax(rp, fabs(c[le] - z[le]));
          nA = fmax(nA, fmax(fabs(c[le]), fabs(z[le])));
        }
      }
    }
    rp = quad_max(rp); rd = quad_max(rd); nA = quad_max(nA); nD = quad_max(nD);
    const bool conv = (rp <= P.eps_abs + P.eps_rel * nA) && (rd <= P.eps_abs + P.eps_rel * fmax(nD, qn));
    bool pinf = false;
    if (HAS_G) {  // OSQP primal infeasibility certificate on delta_y of the last iteration
      double ndy = 0.0, supp = 0.0;
#pragma unroll
      for (int le = 0; le < EPL; le++) {
        const int e = 8 * (le >> 1) + 2 * l4 + (le & 1);
        const double dy = sRho[e] * dys[le];
        dys[le] = dy;
        const double off = ((boxbits >> le) & 1u) ? 0.0 : x[le];
        ndy = fmax(ndy, fabs(dy));
        supp += (sHi[e] + off) * fmax(dy, 0.0) + (sLo[e] + off) * fmin(dy, 0.0);
      }
      ndy = quad_max(ndy); supp = quad_sum(supp);
      const bool cand = (pi >= 0) && !conv && (ndy > P.eps_pinf) && (supp < -P.eps_pinf * ndy);
      if (__any_sync(0xffffffffu, cand)) {
        double in[EPL], c[EPL];
#pragma unroll
        for (int le = 0; le < EPL; le++) in[le] = ((boxbits >> le) & 1u) ? 0.0 : dys[le];
        mma_pass(sC, in, c);
        double atdy = 0.0;
#pragma unroll
        for (int le = 0; le < EPL; le++)
          if ((boxbits >> le) & 1u) atdy = fmax(atdy, fabs(c[le] + dys[le]));
        atdy = quad_max(atdy);
        pinf = cand && (atdy <= P.eps_pinf * ndy);
      }
    }
    const bool fin = (pi >= 0) && (conv || pinf || it_s >= max_iter);
    if (fin) {
#pragma unroll
      for (int tn = 0; tn < NTL; tn++) {
        const int e = 8 * tn + 2 * l4;
        if (((nz & 1) == 0) && e + 1 < nz) {
          *reinterpret_cast<double2*>(&P.v_out[pi * nz + e]) = make_double2(x[2 * tn], x[2 * tn + 1]);
        } else {
          if (e < nz) P.v_out[pi * nz + e] = x[2 * tn];
          if (e + 1 < nz) P.v_out[pi * nz + e + 1] = x[2 * tn + 1];
        }
        if (P.y_out != nullptr) {
#pragma unroll
          for (int jj = 0; jj < 2; jj++)
            if (e + jj < P.nt) P.y_out[pi * P.nt + e + jj] = (HAS_G ? sRho[e + jj] : rho_s) * ys[2 * tn + jj];
        }
      }
      if (l4 == 0) {
        P.status[pi] = conv ? 1 : (pinf ? -3 : -2);
        P.iters[pi] = it_s;
        P.pres[pi] = rp;
        P.dres[pi] = rd;
      }
      pi = -1;
    }
  }
}

}  // namespace mpcb
