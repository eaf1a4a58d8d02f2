// On-chip batched OSQP-style ADMM for the condensed linear-MPC QP (sm_100a).
//
// Replaces the per-problem `JuMP.optimize!` -> OSQP call (/root/reference/src/main/computation_mpc.jl:41) for
// thousands of problems that share one designed controller.  One warp owns 8 problem SLOTS (one per lane quad).
// Per ADMM iteration the multi-RHS contraction  [x~; z~_g] = T r  (T = [I;G] K^-1 [I,G'], cached per system) runs on
// the FP64 tensor path as DMMA.8x8x4 with problems along M, the fragment-ordered T streamed from shared memory as
// the B operand and the per-problem right-hand side produced in registers as the A operand.  Because T is symmetric
// and we control the row order of the B fragments, the C fragment of one iteration IS the A fragment layout of the
// next: the ADMM state never leaves the register file between iterations, and projection / dual update / residual
// reductions are fused right behind the MMAs (quad shuffles only).
// Slots that terminate are written out and refilled from a global work counter at every check point, so iteration
// count variance across the batch does not idle the FP64 pipe.
//
// On B200 DMMA and scalar FP64 instructions share one pipe (ncu: math-pipe-throttle on DFMA while DMMA runs; see
// profiles/), so the elementwise step is written for the fewest FP64 instructions.  Box-only problems (mg = 0) keep
//     c = (1-alpha) z + y/rho        r = rho (z - y/rho) + sigma x - q   (the next MMA operand)
// per row, so that one iteration is  w = alpha t + c;  z+ = clamp(w);  c+ = w - alpha z+;  r+ = rho (2 z+ - w) + ...
// (6 FP64 instructions per row with sigma = 0, 9 with sigma > 0).  Termination is evaluated at (x~, z+, y+): the dual
// residual of the x-subproblem minimiser is available in closed form from the cached factor,
//     Pc x~ = r - (sigma + rho) x~     (mg = 0),
// so box-only problems need no second operator pass; problems with general rows run one pass with C = [[Pc,G'],[G,0]].
//
// Row -> lane mapping: lane = 4*g + l4 (g = slot 0..7, l4 = 0..3); local row le = 2*t + j  <->  row e = 8*t + 2*l4 + j.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mpcb200.h"

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
  int nball;            // contractive terminal set: general rows nz .. nz+nball-1 form one ball (HAS_G kernels only)
  double rho_box, sigma, alpha, eps_abs, eps_rel, eps_pinf;
  int max_iter, check_every;
  // batch
  long long batch;
  const double* x0;     // [batch][nx]
  const double* xref;   // [batch or 1][nx]
  const double* uref;   // [batch or 1][nu]
  int xref_bc, uref_bc;
  const double* Lv;     // [np][NT]  transposed cold-start map (rows < nz: v_unc = Lv p; zero beyond) or null (settings.cold_init = 0)
  const double* warm_v; // [batch][nz] or null
  const double* warm_y; // [batch][nt] or null
  double* v_out;        // [batch][nz]  absolute inputs
  double* y_out;        // [batch][nt] or null
  int32_t* status;
  int32_t* iters;
  double* pres;
  double* dres;
  // second pass over the problems the first pass left unsolved (rho ladder, mpcb_api.cu): ticket t works on problem remap[t],
  // the number of tickets is read from device memory, and the reported iteration count continues from the first pass
  const int32_t* remap;                  // null: ticket = problem index
  const unsigned long long* batch_dev;   // null: `batch` tickets
  int iters_add;
  // division of a second pass between the slot kernels and the CTA-cooperative kernel (admm_coop.cuh): with a device-side ticket count, a slot
  // kernel exits at once when the count is <= tickets_skip_le (the cooperative kernel took them), the cooperative kernel when it is > tickets_max
  long long tickets_skip_le;             // 0 (zero-initialised): never skips real work
  long long tickets_max;                 // cooperative kernel only; < 0: no limit
  unsigned long long* counter;  // [0] work queue head, [1] CTAs that have drained it; both zero at launch, the last CTA to
                                // finish re-zeroes them so that back-to-back launches need no memset in between
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// max / clamp without fmax's NaN plumbing (inputs are finite): one DSETP + two FSEL each
__device__ __forceinline__ double dmaxf(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double dclamp(double w, double lo, double hi) {
  const double t = w < lo ? lo : w;
  return t > hi ? hi : t;
}
// The FP64 pipe is the bound of these kernels (DMMA and scalar FP64 share it), the integer pipe idles: comparisons of doubles
// run on the integer pipe instead.  dkey maps a double to a 64-bit integer with the same ordering (sign-magnitude -> two's
// complement; -0.0 sorts just below +0.0, NaNs beyond the infinities), so a clamp is two integer compares + selects, and the
// max of ABSOLUTE values is an unsigned integer max of the bit patterns with the sign cleared.  Values are never changed.
__device__ __forceinline__ long long dkey(double x) {
  const long long b = __double_as_longlong(x);
  return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}
#ifndef MPCB_FUSED_ED
#define MPCB_FUSED_ED 0     // 1: software-pipelined update+product passes (bit-identical; measured no faster: 0.659 vs 0.671 of peak on the bench batch)
#endif
#ifndef MPCB_INT_CLAMP
#define MPCB_INT_CLAMP 0    // measured on B200: the integer clamp is SLOWER (0.515 vs 0.492 ms, QT H=20) -- longer dependent chains per row; kept as a knob
#endif
#ifndef MPCB_INT_MAX
#define MPCB_INT_MAX 1      // 0: A/B knob, the residual maxima back on the FP64 pipe
#endif
__device__ __forceinline__ double iclamp_k(double w, double lo, double hi, long long klo, long long khi) {
  if (!MPCB_INT_CLAMP) { const double t0 = w < lo ? lo : w; return t0 > hi ? hi : t0; }
  const long long kw = dkey(w);
  const double t = kw < klo ? lo : w;
  return kw > khi ? hi : t;
}
__device__ __forceinline__ unsigned long long absbits(double x) {      // inline PTX: written in C++, the mask is recognised as fabs and comes back as a DADD
  if (!MPCB_INT_MAX) return (unsigned long long)__double_as_longlong(fabs(x));
  unsigned long long r;
  asm("{ .reg .b32 lo, hi; mov.b64 {lo, hi}, %1; and.b32 hi, hi, 0x7fffffff; mov.b64 %0, {lo, hi}; }" : "=l"(r) : "d"(x));
  return r;
}
__device__ __forceinline__ unsigned long long umax64(unsigned long long a, unsigned long long b) {
  if (!MPCB_INT_MAX) return (unsigned long long)__double_as_longlong(dmaxf(__longlong_as_double((long long)a), __longlong_as_double((long long)b)));
  return a > b ? a : b;
}
__device__ __forceinline__ double quad_max(double v) {
  v = dmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = dmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}
__device__ __forceinline__ double quad_sum(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

#ifndef MPCB_R_SMEM
#define MPCB_R_SMEM 0
#endif
constexpr bool R_SMEM = MPCB_R_SMEM != 0;
constexpr int ONCHIP_THREADS = 128;
constexpr int ONCHIP_WARPS = ONCHIP_THREADS / 32;

// per-warp [row][lane] slices of the box-only path: x~ and y+ of the checking iteration, plus the operand r when one of the
// experiment knobs that stage it is on
constexpr int ONCHIP_SLICES = (MPCB_R_SMEM != 0 || MPCB_FUSED_ED != 0) ? 3 : 2;

// shared memory: T, C fragments, Lt, lo, hi, rho, rinv, then per-warp parameter staging [8][npad]
__host__ __device__ inline size_t onchip_smem_bytes(int NT, int np, bool has_g) {
  int npad = (np + 1) & ~1;
  return sizeof(double) * ((size_t)(has_g ? 2 : 1) * NT * NT + (size_t)np * NT + 4 * NT + (size_t)ONCHIP_WARPS * 8 * npad +
                           (has_g ? 0 : (size_t)ONCHIP_WARPS * (NT / 4) * 32 * ONCHIP_SLICES));
}

// NT: padded operator size; HAS_G: general rows present; SIG: sigma != 0 (box-only kernels drop the x state otherwise)
template <int NT, bool HAS_G, bool SIG, int MINB>
__global__ void __launch_bounds__(ONCHIP_THREADS, MINB) admm_onchip_kernel(const OnchipParams P) {
  constexpr int EPL = NT / 4;  // rows per lane
  constexpr int KS = NT / 4;   // k-steps
  constexpr int NTL = NT / 8;  // n-tiles
  constexpr bool KEEP_X = HAS_G || SIG;
  extern __shared__ __align__(16) double smem[];
  double* sT = smem;
  double* sC = sT + NT * NT;                       // only present (and only read) when HAS_G
  double* sL = sT + (HAS_G ? 2 : 1) * NT * NT;
  double* sLo = sL + P.np * NT;
  double* sHi = sLo + NT;
  double* sRho = sHi + NT;
  double* sRinv = sRho + NT;
  const int npad = (P.np + 1) & ~1;
  double* sP = sRinv + NT + (threadIdx.x >> 5) * 8 * npad;
  // Experiment knob MPCB_R_SMEM (default off): stage the MMA A operand r through shared memory ([k-step][lane], each
  // lane reads back only what it wrote) so the k-step loop is a REAL loop and the DMMA issue order stays k-step-major
  // (fully unrolled, ptxas chains the DMMAs of one accumulator back to back: 26-cycle dependent latency vs a 16-cycle
  // issue interval).  Measured on B200 (QT, H=20, 65536 problems): 0.531 ms staged vs 0.493 ms with r in registers --
  // the extra LDS/STS traffic costs more than the chaining, which three warps per scheduler already hide.
  double* sXt = sRinv + NT + ONCHIP_WARPS * 8 * npad + (threadIdx.x >> 5) * KS * 32 * ONCHIP_SLICES + (threadIdx.x & 31);   // x~ / y+ of the checking iteration, [row][lane]
  double* sYo = sXt + KS * 32;
  double* sR = sYo + (ONCHIP_SLICES == 3 ? KS * 32 : 0);   // only with a staging knob on (aliases y+ otherwise and is never touched)

  for (int i = threadIdx.x; i < NT * NT; i += ONCHIP_THREADS) {
    sT[i] = P.Tfrag[i];
    if (HAS_G) sC[i] = P.Cfrag[i];
  }
  for (int i = threadIdx.x; i < P.np * NT; i += ONCHIP_THREADS) sL[i] = P.Lt[i];
  for (int i = threadIdx.x; i < NT; i += ONCHIP_THREADS) {
    sLo[i] = P.lo[i]; sHi[i] = P.hi[i];
    if (HAS_G) { sRho[i] = P.rho[i]; sRinv[i] = P.rinv[i]; }
    else { sRho[i] = __longlong_as_double(dkey(P.lo[i])); sRinv[i] = __longlong_as_double(dkey(P.hi[i])); }   // box-only: the slots hold the bounds' integer keys
  }
  __syncthreads();
  const long long* sLoK = reinterpret_cast<const long long*>(sRho);
  const long long* sHiK = reinterpret_cast<const long long*>(sRinv);

  const int lane = threadIdx.x & 31, g = lane >> 2, l4 = lane & 3;
  const double sigma = P.sigma, alpha = P.alpha, oma = 1.0 - P.alpha;
  const double rho_s = P.rho_box, rinv_s = 1.0 / P.rho_box, sig_rho = P.sigma + P.rho_box;
  const int nz = P.nz;

  // Per owned row.  Box-only: c, r, q (+ x when SIG).  With general rows: z, ys (kept in c, r), q (box rows) or the
  // per-problem bound offset b(p) (general rows), x.
  double c[EPL], q[EPL];
  double r[(HAS_G || !R_SMEM) ? EPL : 1];
  double x[KEEP_X ? EPL : 1];
  double dys[HAS_G ? EPL : 1];
  unsigned boxbits = 0;
#pragma unroll
  for (int le = 0; le < EPL; le++) {
    const int e = 8 * (le >> 1) + 2 * l4 + (le & 1);
    if (!HAS_G || e < nz) boxbits |= 1u << le;
    c[le] = q[le] = 0.0;
    if (HAS_G || !R_SMEM) r[le] = 0.0; else sR[le * 32] = 0.0;
    if (KEEP_X) x[le] = 0.0;
    if (HAS_G) dys[le] = 0.0;
  }
  (void)dys; (void)x; (void)r;
  long long pi = -1;   // problem held by this slot (quad-uniform)
  int it_s = 0;        // its iteration count
  double qn = 0.0;     // ||q||_inf
  double rad = 0.0;    // contractive terminal set: radius sqrt(0.9) |x0 - xref|_2 of this slot's ball
  bool exhausted = false;
  const int max_iter = ((P.max_iter + P.check_every - 1) / P.check_every) * P.check_every;
  long long batch_eff = P.batch_dev ? (long long)*P.batch_dev : P.batch;
  if (P.batch_dev && batch_eff <= P.tickets_skip_le) batch_eff = 0;

  // one product  out = M in  with M in fragment order
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
      if (HAS_G && P.nball > 0) {
        double d2 = 0.0;
        if (fresh)
          for (int j = l4; j < P.nx; j += 4) { const double dv = sP[g * npad + j] - sP[g * npad + P.nx + j]; d2 = fma(dv, dv, d2); }
        d2 = quad_sum(d2);
        if (fresh) rad = sqrt(0.9 * d2);
      }
      double zw[HAS_G ? EPL : 1];   // warm-start z of general rows needs G x0 (one C pass)
      (void)zw;
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
          q[le] = acc[le];               // box rows: q ; general rows: bound offset b(p)
          if ((boxbits >> le) & 1u) m = dmaxf(m, fabs(acc[le]));
          acc[le] = 0.0;
        }
        const bool cold_pt = P.Lv != nullptr && P.warm_v == nullptr;     // settings.cold_init: start at the clipped unconstrained optimum Lv p
        if (cold_pt) {
          for (int j = 0; j < P.np; j++) {
            const double pj = sP[g * npad + j];
#pragma unroll
            for (int t = 0; t < NTL; t++) {
              const double2 l2 = __ldg(reinterpret_cast<const double2*>(&P.Lv[j * NT + 8 * t + 2 * l4]));
              acc[2 * t] = fma(l2.x, pj, acc[2 * t]);
              acc[2 * t + 1] = fma(l2.y, pj, acc[2 * t + 1]);
            }
          }
        }
#pragma unroll
        for (int le = 0; le < EPL; le++) {
          const int e = 8 * (le >> 1) + 2 * l4 + (le & 1);
          const bool box = (boxbits >> le) & 1u;
          double v0 = 0.0, ys0 = 0.0;    // OSQP warm start: x = v0, z = A x, y = y0
          if (P.warm_v != nullptr) {
            if (box && e < nz) v0 = P.warm_v[pi * nz + e];
            if (e < P.nt) ys0 = P.warm_y[pi * P.nt + e] * (HAS_G ? sRinv[e] : rinv_s);
          } else if (cold_pt && box && e < nz) {
            v0 = dclamp(acc[le], sLo[e], sHi[e]);
            ys0 = -MPCB_INIT_KAPPA * (v0 - acc[le]);      // y / rho with y = -kappa rho (x - v_unc)
          }
          if (KEEP_X) x[le] = box ? v0 : 0.0;
          if (!HAS_G) {
            c[le] = fma(oma, v0, ys0);
            const double r0 = fma(rho_s, v0 - ys0, fma(sigma, v0, -q[le]));
            if (R_SMEM) sR[le * 32] = r0; else r[le] = r0;
          } else {
            c[le] = v0;    // z
            r[le] = ys0;   // ys
          }
        }
        qn = m;
      }
      qn = quad_max(qn);
      if (HAS_G && (P.warm_v != nullptr || P.Lv != nullptr)) {
        double in[EPL], out[EPL];
#pragma unroll
        for (int le = 0; le < EPL; le++) in[le] = (fresh && ((boxbits >> le) & 1u)) ? c[le] : 0.0;
        mma_pass(sC, in, out);
#pragma unroll
        for (int le = 0; le < EPL; le++)
          if (fresh && !((boxbits >> le) & 1u)) c[le] = out[le];
      }
    }
    if (!__any_sync(0xffffffffu, pi >= 0)) break;

    // ------------------------------------------------------------------ check_every ADMM iterations, the last one checks
    double rp = 0.0, rd = 0.0, nA = 0.0, nD = 0.0;
    unsigned long long urp = 0ULL, urd = 0ULL, unA = 0ULL, unD = 0ULL;     // box-only path: the same maxima as bit patterns (integer pipe)
    double xt[HAS_G ? EPL : 1], yo[HAS_G ? EPL : 1];   // x~ (general rows: z~) and y+ of the checking iteration
    (void)xt; (void)yo;
    if (!HAS_G && MPCB_FUSED_ED && !R_SMEM) {
      // -------- box-only fast path, software-pipelined: the elementwise update of row s (which yields the A operand of k-step s of
      // the NEXT product) is issued right before that k-step's DMMAs, so one warp keeps the tensor pipe busy through what used to
      // be its elementwise phase instead of relying on the other warps of the scheduler.  One period of check_every iterations =
      // [plain product] [check_every - 1 fused update+product passes] [checking update].  Same operations per row in the same
      // order as the plain loop: results are bit-identical.  The operand r of the last product is parked in shared memory for
      // the dual residual (its registers are reused by the second accumulator set during the fused passes).
      double t[EPL];
#pragma unroll
      for (int i = 0; i < EPL; i++) t[i] = 0.0;
#pragma unroll
      for (int s = 0; s < KS; s++) {
#pragma unroll
        for (int tn = 0; tn < NTL; tn++) dmma884(t[2 * tn], t[2 * tn + 1], r[s], sT[(s * NTL + tn) * 32 + lane]);
      }
      if (P.check_every == 1) {
#pragma unroll
        for (int s = 0; s < KS; s++) sR[s * 32] = r[s];
      }
      for (int ii = 1; ii < P.check_every; ii++) {
        const bool last = (ii == P.check_every - 1);
        double t2[EPL];
#pragma unroll
        for (int i = 0; i < EPL; i++) t2[i] = 0.0;
#pragma unroll
        for (int s = 0; s < KS; s++) {
          const int e = 8 * (s >> 1) + 2 * l4 + (s & 1);
          const double w = fma(alpha, t[s], c[s]);
          const double zn = dclamp(w, sLo[e], sHi[e]);
          c[s] = fma(-alpha, zn, w);
          const double d = fma(2.0, zn, -w);
          double rn;
          if (SIG) {
            x[s] = fma(alpha, t[s], oma * x[s]);
            rn = fma(rho_s, d, fma(sigma, x[s], -q[s]));
          } else {
            rn = fma(rho_s, d, -q[s]);
          }
          if (last) sR[s * 32] = rn;
#pragma unroll
          for (int tn = 0; tn < NTL; tn++) dmma884(t2[2 * tn], t2[2 * tn + 1], rn, sT[(s * NTL + tn) * 32 + lane]);
        }
#pragma unroll
        for (int i = 0; i < EPL; i++) t[i] = t2[i];
      }
      // checking update: residuals of (x~, z+, y+) with Pc x~ = r - (sigma + rho) x~, and the operand of the next period
#pragma unroll
      for (int tn = 0; tn < NTL; tn++) {
        const double2 lo2 = *reinterpret_cast<const double2*>(&sLo[8 * tn + 2 * l4]);
        const double2 hi2 = *reinterpret_cast<const double2*>(&sHi[8 * tn + 2 * l4]);
#pragma unroll
        for (int jj = 0; jj < 2; jj++) {
          const int le = 2 * tn + jj;
          const double w = fma(alpha, t[le], c[le]);
          const double zn = dclamp(w, jj ? lo2.y : lo2.x, jj ? hi2.y : hi2.x);
          const double pc = fma(-sig_rho, t[le], sR[le * 32]);
          const double yb = rho_s * (w - zn);
          urp = umax64(urp, absbits(t[le] - zn));
          urd = umax64(urd, absbits(pc + q[le] + yb));
          unA = umax64(unA, umax64(absbits(t[le]), absbits(zn)));
          unD = umax64(unD, umax64(absbits(pc), absbits(yb)));
          sXt[le * 32] = t[le];
          if (P.y_out != nullptr) sYo[le * 32] = yb;
          c[le] = fma(-alpha, zn, w);
          const double d = fma(2.0, zn, -w);
          if (SIG) {
            x[le] = fma(alpha, t[le], oma * x[le]);
            r[le] = fma(rho_s, d, fma(sigma, x[le], -q[le]));
          } else {
            r[le] = fma(rho_s, d, -q[le]);
          }
        }
      }
    } else
    for (int ii = 0; ii < P.check_every; ii++) {
      const bool chk = (ii == P.check_every - 1);
      double t[EPL];
#pragma unroll
      for (int i = 0; i < EPL; i++) t[i] = 0.0;
      if (!HAS_G) {
        // -------- box-only fast path: r (staged in shared memory) is the MMA operand
        if (R_SMEM) {
          const double* tp = sT + lane;
#pragma unroll 1
          for (int s = 0; s < KS; s++) {
            const double a = sR[s * 32];
#pragma unroll
            for (int tn = 0; tn < NTL; tn++) dmma884(t[2 * tn], t[2 * tn + 1], a, tp[(s * NTL + tn) * 32]);
          }
        } else {
#pragma unroll
          for (int s = 0; s < KS; s++) {
#pragma unroll
            for (int tn = 0; tn < NTL; tn++) dmma884(t[2 * tn], t[2 * tn + 1], r[s], sT[(s * NTL + tn) * 32 + lane]);
          }
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
            const double w = fma(alpha, t[le], c[le]);
            const double zn = iclamp_k(w, jj ? lo2.y : lo2.x, jj ? hi2.y : hi2.x, jj ? kl2.y : kl2.x, jj ? kh2.y : kh2.x);
            if (chk) {   // residuals of (x~, z+, y+): Pc x~ = r - (sigma + rho) x~
              const double pc = fma(-sig_rho, t[le], R_SMEM ? sR[le * 32] : r[le]);
              const double yb = rho_s * (w - zn);
              urp = umax64(urp, absbits(t[le] - zn));
              urd = umax64(urd, absbits(pc + q[le] + yb));
              unA = umax64(unA, umax64(absbits(t[le]), absbits(zn)));
              unD = umax64(unD, umax64(absbits(pc), absbits(yb)));
              sXt[le * 32] = t[le];
              if (P.y_out != nullptr) sYo[le * 32] = yb;
            }
            c[le] = fma(-alpha, zn, w);
            const double d = fma(2.0, zn, -w);
            if (SIG) {
              x[le] = fma(alpha, t[le], oma * x[le]);
              const double rn = fma(rho_s, d, fma(sigma, x[le], -q[le]));
              if (R_SMEM) sR[le * 32] = rn; else r[le] = rn;
            } else {
              const double rn = fma(rho_s, d, -q[le]);
              if (R_SMEM) sR[le * 32] = rn; else r[le] = rn;
            }
          }
        }
      } else {
        // -------- general rows present: c = z, r = ys, q = q (box) / b (general), per-row rho
#pragma unroll
        for (int s = 0; s < KS; s++) {
          const int e = 8 * (s >> 1) + 2 * l4 + (s & 1);
          double rr = sRho[e] * (c[s] - r[s]);
          if ((boxbits >> s) & 1u) rr += fma(sigma, x[s], -q[s]);
#pragma unroll
          for (int tn = 0; tn < NTL; tn++) dmma884(t[2 * tn], t[2 * tn + 1], rr, sT[(s * NTL + tn) * 32 + lane]);
        }
        // contractive terminal set (design_mpc.jl:333-340): the rows nz .. nz+nball-1 of a slot are projected onto ONE ball
        // |w - b|_2 <= rad (centre b = the per-problem offset kept in q) instead of a box; the squared distance is summed over
        // the quad that owns the slot
        double bscale = 1.0;
        if (P.nball > 0) {
          double d2 = 0.0;
#pragma unroll
          for (int le = 0; le < EPL; le++) {
            const int e = 8 * (le >> 1) + 2 * l4 + (le & 1);
            if (e >= nz && e < nz + P.nball) { const double dv = fma(oma, c[le], alpha * t[le]) + r[le] - q[le]; d2 = fma(dv, dv, d2); }
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
            const int le = 2 * tn + jj;
            const bool box = (boxbits >> le) & 1u;
            const double at = alpha * t[le];
            const double w = fma(oma, c[le], at) + r[le];
            double lo_e = jj ? lo2.y : lo2.x, hi_e = jj ? hi2.y : hi2.x;
            if (box) x[le] = fma(oma, x[le], at);
            else { lo_e += q[le]; hi_e += q[le]; }
            double zn = dclamp(w, lo_e, hi_e);
            if (P.nball > 0) {
              const int e = 8 * tn + 2 * l4 + jj;
              if (e >= nz && e < nz + P.nball) zn = fma(w - q[le], bscale, q[le]);
            }
            const double yn = w - zn;
            if (chk) {
              const double rho_e = jj ? rh2.y : rh2.x;
              dys[le] = rho_e * (yn - r[le]);
              rp = dmaxf(rp, fabs(t[le] - zn));                    // A x~ = [x~; G x~] = t
              nA = dmaxf(nA, dmaxf(fabs(t[le]), fabs(zn)));
              xt[le] = t[le];
              yo[le] = rho_e * yn;
            }
            c[le] = zn;
            r[le] = yn;
          }
        }
      }
    }
    it_s += P.check_every;

    // ------------------------------------------------------------------ termination (OSQP criteria at x~, z+, y+)
    // With general rows the dual residual needs a pass with C; a slot can only terminate on it when its primal residual has converged (or at
    // the cap, or with an infeasibility candidate, whose reported residuals must be complete): while no slot of the warp is there, it is skipped.
    bool pinf = false;
    bool prim_ok = true, cand0 = false, want_rd = true;
    double ndy = 0.0, supp = 0.0;
    if (HAS_G) {
      rp = quad_max(rp); nA = quad_max(nA);
#pragma unroll
      for (int le = 0; le < EPL; le++) {
        const int e = 8 * (le >> 1) + 2 * l4 + (le & 1);
        const double dy = dys[le];
        const double off = ((boxbits >> le) & 1u) ? 0.0 : q[le];
        ndy = dmaxf(ndy, fabs(dy));
        supp += (sHi[e] + off) * dmaxf(dy, 0.0) + (sLo[e] + off) * (dy < 0.0 ? dy : 0.0);
      }
      ndy = quad_max(ndy); supp = quad_sum(supp);
      prim_ok = rp <= P.eps_abs + P.eps_rel * nA;
      cand0 = (pi >= 0) && (P.nball == 0) && (ndy > P.eps_pinf) && (supp < -P.eps_pinf * ndy);   // no certificate is evaluated for ball rows
      want_rd = (pi >= 0) && (prim_ok || cand0 || it_s >= max_iter);
    }
    if (HAS_G && __any_sync(0xffffffffu, want_rd)) {
      double in[EPL], cc[EPL];
#pragma unroll
      for (int le = 0; le < EPL; le++) in[le] = ((boxbits >> le) & 1u) ? xt[le] : yo[le];   // [x~; y_g]
      mma_pass(sC, in, cc);                                                                 // [Pc x~ + G' y_g ; G x~]
#pragma unroll
      for (int le = 0; le < EPL; le++)
        if ((boxbits >> le) & 1u) {
          rd = dmaxf(rd, fabs(cc[le] + q[le] + yo[le]));
          nD = dmaxf(nD, dmaxf(fabs(cc[le]), fabs(yo[le])));
        }
    }
    if (!HAS_G) { rp = __longlong_as_double((long long)urp); rd = __longlong_as_double((long long)urd); nA = __longlong_as_double((long long)unA); nD = __longlong_as_double((long long)unD); }
    if (!HAS_G) { rp = quad_max(rp); nA = quad_max(nA); }
    rd = quad_max(rd); nD = quad_max(nD);
    const bool conv = want_rd && (rp <= P.eps_abs + P.eps_rel * nA) && (rd <= P.eps_abs + P.eps_rel * dmaxf(nD, qn));
    if (HAS_G) {  // OSQP primal infeasibility certificate on delta_y of the last iteration
      const bool cand = cand0 && !conv;
      if (__any_sync(0xffffffffu, cand)) {
        double in[EPL], cc[EPL];
#pragma unroll
        for (int le = 0; le < EPL; le++) in[le] = ((boxbits >> le) & 1u) ? 0.0 : dys[le];
        mma_pass(sC, in, cc);
        double atdy = 0.0;
#pragma unroll
        for (int le = 0; le < EPL; le++)
          if ((boxbits >> le) & 1u) atdy = dmaxf(atdy, fabs(cc[le] + dys[le]));
        atdy = quad_max(atdy);
        pinf = cand && (atdy <= P.eps_pinf * ndy);
      }
    }
    const bool fin = (pi >= 0) && (conv || pinf || it_s >= max_iter);
    if (fin) {
#pragma unroll
      for (int tn = 0; tn < NTL; tn++) {
        const int e = 8 * tn + 2 * l4;
        const double x0v = HAS_G ? xt[2 * tn] : sXt[(2 * tn) * 32], x1v = HAS_G ? xt[2 * tn + 1] : sXt[(2 * tn + 1) * 32];
        if (((nz & 1) == 0) && e + 1 < nz) {
          *reinterpret_cast<double2*>(&P.v_out[pi * nz + e]) = make_double2(x0v, x1v);
        } else {
          if (e < nz) P.v_out[pi * nz + e] = x0v;
          if (e + 1 < nz) P.v_out[pi * nz + e + 1] = x1v;
        }
        if (P.y_out != nullptr) {
#pragma unroll
          for (int jj = 0; jj < 2; jj++)
            if (e + jj < P.nt) P.y_out[pi * P.nt + e + jj] = HAS_G ? yo[2 * tn + jj] : sYo[(2 * tn + jj) * 32];
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
  // every warp of this CTA has seen the queue empty: the last CTA of the grid re-arms the queue for the next launch
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long prev = atomicAdd(P.counter + 1, 1ULL);
    if (prev == (unsigned long long)gridDim.x - 1ULL) { P.counter[0] = 0ULL; P.counter[1] = 0ULL; }
  }
}

}  // namespace mpcb
