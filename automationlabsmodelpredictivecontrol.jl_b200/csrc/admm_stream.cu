// Streamed batched ADMM for operators that do not fit the register file (nt = nz + mg > 64), sm_100a.
//
// Same algorithm and termination as the on-chip kernel (admm_onchip.cuh), but the ADMM state
// lives in HBM/L2 as problem-major matrices [rows][NTp] and every iteration is ONE kernel: a tiled FP64 tensor GEMM
//     OUT[b][:] = R[b][:] * T          (T symmetric, cached per system, L2 resident)
// with the ADMM elementwise step (relaxation, box projection, dual update, next right-hand side, residual partials)
// fused into the epilogue, so the state is read and written exactly once per iteration.  Operand tiles are staged into
// padded shared-memory rows with 16-byte cp.async copies, STAGES deep; the MMA is DMMA.8x8x4 (FP64 has no tcgen05 kind)
// with 32x32 per-warp register tiles.
// Terminated problems are written out and the active rows are compacted so the GEMM shrinks with the batch.
#include "admm_stream.cuh"

#include <algorithm>
#include <cstdio>
#include <vector>

namespace mpcb {

namespace {

constexpr int BM = 128, BN = 64, BK = 16, LDS_ = 20, STAGES = 3, THREADS = 256;
constexpr int STAGE_DOUBLES = (BM + BN) * LDS_;
constexpr size_t SMEM_BYTES = sizeof(double) * STAGE_DOUBLES * STAGES;

enum : int { MODE_NORMAL = 0, MODE_SAVE_YP = 1, MODE_CHECK = 2 };

struct IterParams {
  double* R0;          // [rows][NTp] MMA operand of the period's first iteration (iterations alternate R0 -> R1 -> R0 ...)
  double* R1;
  int iters;           // iterations fused into this launch (= check_every)
  int* sched;          // device: [0] tile ticket counter, [1 + it * nrb + rb] column tiles of (iteration it, row block rb) finished
  int nrb, ncb;
  const double* T;     // [NTp][NTp]
  double* Cst;         // c = (1-alpha) z + y/rho
  const double* QB;    // box cols: q ; general cols: bound offset b(p)
  double* X;           // relaxed x (only touched when sigma != 0)
  const double* lo; const double* hi; const double* rho;   // [NTp]
  double* XT; double* YO; double* YP;                      // candidate x~ / y+ (check), y of the iteration before (certificate)
  unsigned long long* red;                                 // [rows][4]: rp, rd, nA, nD as bit patterns of non-negative doubles
  int rows, NTp, nz, nt, mg;
  double alpha, sigma;
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double dmaxf(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double dclamp(double w, double lo, double hi) { const double t = w < lo ? lo : w; return t > hi ? hi : t; }
__device__ __forceinline__ void atomic_max_nn(unsigned long long* addr, double v) {   // v >= 0: bit patterns order like the values
  atomicMax(addr, (unsigned long long)__double_as_longlong(v));
}

// Tiled GEMM core: acc[mt][nt2][2] for the warp's 32x32 tile of OUT = In * M (M symmetric, row-major == column-major).
// Operand tiles [rows][BK] are staged with 16-byte cp.async (LDGSTS, L2 -> shared memory without a register round trip)
// into rows padded to LDS_ doubles, so that the 64-bit fragment loads of every half-warp hit 32 distinct banks.  One
// __syncthreads per k-tile: it publishes the tile that just landed and retires the stage the next prefetch overwrites.
// (History: the first version staged these tiles with per-row cp.async.bulk + mbarrier copies of 128 bytes; ncu showed the
// DMMA pipe 17 % active with the warps parked on the barrier behind the TMA queue -- 192 tiny bulk copies per k-tile --
// see profiles/r01/prof_stream_lti_r01a.txt.  FP64 mma.sync fragments need padded rows, which the tensor-map TMA cannot
// produce, so the copy engine of choice here is LDGSTS.)
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// kt0: first k-tile of the product (operands known to be zero in columns < kt0 * BK are skipped -- the certificate pass G' dy_g only has
// the mg general columns)
__device__ __forceinline__ void gemm_tile(const double* __restrict__ In, const double* __restrict__ M, int rows, int NTp, int bm0, int bn0,
                                          double (&acc)[4][4][2], double* smem, int kt0 = 0) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, l4 = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;
  const int KT = NTp / BK;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  // this thread's 16-byte chunks of a stage: (BM + BN) rows x (BK / 2) chunks, THREADS chunks per pass
  constexpr int CPR = BK / 2;                                  // chunks per row
  constexpr int PASSES = (BM + BN) * CPR / THREADS;
  static_assert((BM + BN) * CPR % THREADS == 0, "tile must divide evenly over the CTA");
  const double* src[PASSES];
  int dst[PASSES];
#pragma unroll
  for (int q = 0; q < PASSES; q++) {
    const int c = tid + q * THREADS, r = c / CPR, cc = c % CPR;
    if (r < BM) src[q] = In + (size_t)min(bm0 + r, rows - 1) * NTp + cc * 2;      // tail tile: re-read the last row, masked in the epilogue
    else src[q] = M + (size_t)(bn0 + r - BM) * NTp + cc * 2;
    dst[q] = r * LDS_ + cc * 2;
  }
  auto issue = [&](int kt) {
    double* st = smem + (kt % STAGES) * STAGE_DOUBLES;
#pragma unroll
    for (int q = 0; q < PASSES; q++) cp_async16(st + dst[q], src[q] + kt * BK);
  };
#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) {
    if (kt0 + s < KT) issue(kt0 + s);
    cp_async_commit();
  }
  for (int kt = kt0; kt < KT; kt++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (kt + STAGES - 1 < KT) issue(kt + STAGES - 1);
    cp_async_commit();
    const double* sA = smem + (kt % STAGES) * STAGE_DOUBLES + (wm * 32 + g) * LDS_ + l4;
    const double* sB = smem + (kt % STAGES) * STAGE_DOUBLES + BM * LDS_ + (wn * 32 + g) * LDS_ + l4;
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; k4++) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = sA[i * 8 * LDS_ + k4 * 4];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = sB[j * 8 * LDS_ + k4 * 4];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  cp_async_wait<0>();
}

// One launch = `iters` ADMM iterations.  A row block of iteration it+1 needs only the SAME row block of iteration it (all
// of its column tiles), so there is no grid-wide barrier between iterations: CTAs draw (iteration, row block, column tile)
// tickets in that order from a global counter and wait on a per-(iteration, row block) completion count.  This removes the
// launch per iteration and, above all, the idle tail of every iteration's last partial wave (LTI64: 896 tiles on 296 CTA
// slots = 3.03 waves, i.e. a quarter of the machine-time of each iteration was spent waiting for 8 tiles).
// Deadlock-free: tickets are drawn in dependency order and every CTA that holds a ticket is resident.
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(THREADS, 2) stream_iter_kernel(const IterParams P) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_ticket;
  const int tiles_per_iter = P.nrb * P.ncb, total = tiles_per_iter * P.iters;
  for (;;) {
  __syncthreads();                                   // the previous tile's epilogue no longer needs s_ticket / the pipeline buffers
  if (threadIdx.x == 0) s_ticket = atomicAdd(P.sched, 1);
  __syncthreads();
  const int ticket = s_ticket;
  if (ticket >= total) break;
  const int it = ticket / tiles_per_iter, rem = ticket - it * tiles_per_iter, rb = rem / P.ncb, cb = rem - rb * P.ncb;
  if (it > 0) {
    if (threadIdx.x == 0) {
      const int* flag = P.sched + 1 + (it - 1) * P.nrb + rb;
      while (ld_acquire(flag) < P.ncb) __nanosleep(64);
    }
    __syncthreads();
  }
  const double* Rin = (it & 1) ? P.R1 : P.R0;
  double* Rout = (it & 1) ? P.R0 : P.R1;
  const int mode = (it == P.iters - 1) ? MODE_CHECK : ((P.mg > 0 && it == P.iters - 2) ? MODE_SAVE_YP : MODE_NORMAL);
  const int bn0 = cb * BN, bm0 = rb * BM;
  double acc[4][4][2];
  gemm_tile(Rin, P.T, P.rows, P.NTp, bm0, bn0, acc, smem);

  // ---- fused ADMM step on the accumulator fragments: row b = problem, columns n, n+1
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, l4 = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;
  const double alpha = P.alpha, oma = 1.0 - P.alpha, sigma = P.sigma;
  const bool sig = sigma != 0.0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int braw = bm0 + wm * 32 + i * 8 + g;
    const bool valid = braw < P.rows;          // tail tile: compute on the (re-read) last row, store nothing
    const int b = valid ? braw : P.rows - 1;
    double rp = 0.0, rd = 0.0, nA = 0.0, nD = 0.0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int n = bn0 + wn * 32 + j * 8 + 2 * l4;
      const size_t off = (size_t)b * P.NTp + n;
      const double2 c2 = __ldcg(reinterpret_cast<const double2*>(P.Cst + off));      // written by another SM one iteration ago: bypass L1
      const double2 q2 = *reinterpret_cast<const double2*>(P.QB + off);
      const double2 lo2 = *reinterpret_cast<const double2*>(P.lo + n);
      const double2 hi2 = *reinterpret_cast<const double2*>(P.hi + n);
      const double2 rh2 = *reinterpret_cast<const double2*>(P.rho + n);
      double2 x2 = make_double2(0.0, 0.0), rin2 = make_double2(0.0, 0.0);
      if (sig) x2 = __ldcg(reinterpret_cast<const double2*>(P.X + off));
      if (mode == MODE_CHECK && P.mg == 0) rin2 = __ldcg(reinterpret_cast<const double2*>(Rin + off));
      double cn[2], rn[2], xn[2], xt[2], yo[2];
#pragma unroll
      for (int jj = 0; jj < 2; jj++) {
        const bool box = (n + jj) < P.nz;
        const double t = acc[i][j][jj];
        const double cc = jj ? c2.y : c2.x, qq = jj ? q2.y : q2.x, rho_e = jj ? rh2.y : rh2.x;
        double lo_e = jj ? lo2.y : lo2.x, hi_e = jj ? hi2.y : hi2.x;
        if (!box) { lo_e += qq; hi_e += qq; }
        const double w = fma(alpha, t, cc);
        const double zn = dclamp(w, lo_e, hi_e);
        const double yb = rho_e * (w - zn);
        xt[jj] = t; yo[jj] = yb;
        if (mode == MODE_CHECK) {
          rp = dmaxf(rp, fabs(t - zn));
          nA = dmaxf(nA, dmaxf(fabs(t), fabs(zn)));
          if (P.mg == 0) {   // closed-form dual residual: Pc x~ = r - (sigma + rho) x~
            const double pc = fma(-(sigma + rho_e), t, jj ? rin2.y : rin2.x);
            rd = dmaxf(rd, fabs(pc + qq + yb));
            nD = dmaxf(nD, dmaxf(fabs(pc), fabs(yb)));
          }
        }
        cn[jj] = fma(-alpha, zn, w);
        const double d = fma(2.0, zn, -w);
        xn[jj] = 0.0;
        if (box) {
          if (sig) { xn[jj] = fma(alpha, t, oma * (jj ? x2.y : x2.x)); rn[jj] = fma(rho_e, d, fma(sigma, xn[jj], -qq)); }
          else rn[jj] = fma(rho_e, d, -qq);
        } else {
          rn[jj] = rho_e * d;
        }
      }
      if (!valid) continue;
      *reinterpret_cast<double2*>(P.Cst + off) = make_double2(cn[0], cn[1]);
      *reinterpret_cast<double2*>(Rout + off) = make_double2(rn[0], rn[1]);
      if (sig) *reinterpret_cast<double2*>(P.X + off) = make_double2(xn[0], xn[1]);
      if (mode == MODE_SAVE_YP) *reinterpret_cast<double2*>(P.YP + off) = make_double2(yo[0], yo[1]);
      if (mode == MODE_CHECK) {
        *reinterpret_cast<double2*>(P.XT + off) = make_double2(xt[0], xt[1]);
        *reinterpret_cast<double2*>(P.YO + off) = make_double2(yo[0], yo[1]);
      }
    }
    if (mode == MODE_CHECK) {
      // the four lanes of a quad share row b: combine, then one atomic per quad
      for (int o = 1; o <= 2; o <<= 1) {
        rp = dmaxf(rp, __shfl_xor_sync(0xffffffffu, rp, o)); rd = dmaxf(rd, __shfl_xor_sync(0xffffffffu, rd, o));
        nA = dmaxf(nA, __shfl_xor_sync(0xffffffffu, nA, o)); nD = dmaxf(nD, __shfl_xor_sync(0xffffffffu, nD, o));
      }
      if (l4 == 0 && valid) {
        unsigned long long* r4 = P.red + (size_t)b * 4;
        atomic_max_nn(r4 + 0, rp); atomic_max_nn(r4 + 2, nA);
        if (P.mg == 0) { atomic_max_nn(r4 + 1, rd); atomic_max_nn(r4 + 3, nD); }
      }
    }
  }
  // publish this tile: every thread's stores are fenced, then one release increment of the row block's completion count
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(P.sched + 1 + it * P.nrb + rb, 1);
  }   // ticket loop
}

// second operator pass of a check when general rows exist:  [Pc x~ + G' y_g ; G x~] = [x~; y_g] * C
struct CheckParams {
  const double* In;   // [rows][NTp]: box cols x~, general cols y_g
  const double* C;    // [NTp][NTp]
  const double* QB; const double* YO;
  unsigned long long* red;
  double* Out;        // optional: raw product (used for A' dy of the certificate), may be null
  int rows, NTp, nz, mode;   // mode 0: dual residual reductions ; 1: store product only
  int kt0;                   // first k-tile (mode 1: the operand is zero on the box columns)
  // Skipping: a row can only terminate on the dual residual when its primal residual has converged (or at the cap, or as an infeasibility
  // candidate, whose reported residuals must be complete); the certificate product is needed for candidates only.  A CTA whose BM rows hold no
  // such row returns before its GEMM tile (same decisions and reported residuals: stream_decide_kernel tests the primal residual first).
  const int* done; const double* cert;     // cert: [rows][3] ndy, supp, atdy (stream_cert_kernel) or null
  double eps_abs, eps_rel, eps_pinf;
  int want_all;                            // the iteration cap is reached: every active row reports its residuals
};
__global__ void __launch_bounds__(THREADS, 2) stream_check_kernel(const CheckParams P) {
  extern __shared__ __align__(16) double smem[];
  const int bn0 = blockIdx.x * BN, bm0 = blockIdx.y * BM;
  if (!P.want_all) {
    int want = 0;
    for (int r = threadIdx.x; r < BM; r += THREADS) {
      const int b = bm0 + r;
      if (b < P.rows && !P.done[b]) {
        bool cand = false;
        if (P.cert != nullptr) { const double ndy = P.cert[(size_t)b * 3], supp = P.cert[(size_t)b * 3 + 1]; cand = (ndy > P.eps_pinf) && (supp < -P.eps_pinf * ndy); }
        if (P.mode == 1) want |= cand ? 1 : 0;
        else {
          const double rp = __longlong_as_double((long long)P.red[(size_t)b * 4]), nA = __longlong_as_double((long long)P.red[(size_t)b * 4 + 2]);
          want |= (cand || rp <= P.eps_abs + P.eps_rel * nA) ? 1 : 0;
        }
      }
    }
    if (!__syncthreads_or(want)) return;
  }
  double acc[4][4][2];
  gemm_tile(P.In, P.C, P.rows, P.NTp, bm0, bn0, acc, smem, P.kt0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, l4 = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int braw = bm0 + wm * 32 + i * 8 + g;
    const bool valid = braw < P.rows;
    const int b = valid ? braw : P.rows - 1;
    double rd = 0.0, nD = 0.0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int n = bn0 + wn * 32 + j * 8 + 2 * l4;
      const size_t off = (size_t)b * P.NTp + n;
      if (P.mode == 1) { if (valid) *reinterpret_cast<double2*>(P.Out + off) = make_double2(acc[i][j][0], acc[i][j][1]); continue; }
      const double2 q2 = *reinterpret_cast<const double2*>(P.QB + off);
      const double2 y2 = *reinterpret_cast<const double2*>(P.YO + off);
#pragma unroll
      for (int jj = 0; jj < 2; jj++)
        if (n + jj < P.nz) {
          const double gg = acc[i][j][jj], yb = jj ? y2.y : y2.x, qq = jj ? q2.y : q2.x;
          rd = dmaxf(rd, fabs(gg + qq + yb));
          nD = dmaxf(nD, dmaxf(fabs(gg), fabs(yb)));
        }
    }
    if (P.mode == 0) {
      for (int o = 1; o <= 2; o <<= 1) {
        rd = dmaxf(rd, __shfl_xor_sync(0xffffffffu, rd, o)); nD = dmaxf(nD, __shfl_xor_sync(0xffffffffu, nD, o));
      }
      if (l4 == 0 && valid) { atomic_max_nn(P.red + (size_t)b * 4 + 1, rd); atomic_max_nn(P.red + (size_t)b * 4 + 3, nD); }
    }
  }
}

// ---- setup: q = Lq p, b = Lb p, cold-start state -------------------------------------------------------------------
struct InitParams {
  const double* Lt;   // [np][NTp]
  const double *x0, *xref, *uref;
  int xref_bc, uref_bc, nx, nu, np, NTp, nz, nt;
  long long rows;
  double *Cst, *QB, *X, *R, *qn;
  int* idx; int* done; unsigned long long* red;
  // warm start (OSQP: x = v0, z = A x = [v0; G v0], y = y0), both null for a cold start
  const double *warm_v, *warm_y, *C, *rho;     // C: [NTp][NTp] check operator (rows >= nz hold G), rho: [NTp]
  double alpha, sigma;
  const int32_t* remap;                        // row -> problem of the caller's batch (null: identity)
  // settings.cold_init: a cold start begins at x = clip(Lv p), y_box = -kappa rho (x - Lv p), z_g = G x, y_g = 0 (Lv null: zeros)
  const double *Lv, *lo, *hi;                  // Lv: [np][NTp]
};
__global__ void stream_init_kernel(const InitParams P) {
  extern __shared__ double sp[];   // p vector of this problem (+ with cold_init: x [nz], y_box / rho [nz])
  const bool cold_pt = P.Lv != nullptr && P.warm_v == nullptr;
  double* sv = sp + P.np;
  double* sy = sv + P.nz;
  const long long b = blockIdx.x;
  const long long o = P.remap ? (long long)P.remap[b] : b;      // problem of the caller's batch behind this row
  for (int j = threadIdx.x; j < P.np; j += blockDim.x) {
    double v;
    if (j < P.nx) v = P.x0[o * P.nx + j];
    else if (j < 2 * P.nx) v = P.xref[(P.xref_bc ? 0 : o) * P.nx + (j - P.nx)];
    else v = P.uref[(P.uref_bc ? 0 : o) * P.nu + (j - 2 * P.nx)];
    sp[j] = v;
  }
  __syncthreads();
  if (cold_pt) {
    for (int n = threadIdx.x; n < P.nz; n += blockDim.x) {
      double vu = 0.0;
      for (int j = 0; j < P.np; j++) vu = fma(P.Lv[(size_t)j * P.NTp + n], sp[j], vu);
      const double v0 = fmin(fmax(vu, P.lo[n]), P.hi[n]);
      sv[n] = v0; sy[n] = -MPCB_INIT_KAPPA * (v0 - vu);
    }
    __syncthreads();
  }
  double m = 0.0;
  for (int n = threadIdx.x; n < P.NTp; n += blockDim.x) {
    double acc = 0.0;
    for (int j = 0; j < P.np; j++) acc = fma(P.Lt[(size_t)j * P.NTp + n], sp[j], acc);
    const size_t off = (size_t)b * P.NTp + n;
    P.QB[off] = acc;
    if (n < P.nz) m = dmaxf(m, fabs(acc));
    if ((P.warm_v == nullptr && !cold_pt) || n >= P.nt) {
      P.Cst[off] = 0.0; P.X[off] = 0.0;
      P.R[off] = (n < P.nz) ? -acc : 0.0;
    } else {
      const double* v0 = cold_pt ? sv : P.warm_v + (size_t)o * P.nz;
      double z;
      if (n < P.nz) z = v0[n];
      else { z = 0.0; for (int j = 0; j < P.nz; j++) z = fma(P.C[(size_t)n * P.NTp + j], v0[j], z); }     // general row: G v0
      const double rho_n = P.rho[n], ys = cold_pt ? (n < P.nz ? sy[n] : 0.0) : P.warm_y[(size_t)o * P.nt + n] / rho_n;
      P.Cst[off] = fma(1.0 - P.alpha, z, ys);
      P.X[off] = (n < P.nz) ? z : 0.0;
      P.R[off] = rho_n * (z - ys) + ((n < P.nz) ? fma(P.sigma, z, -acc) : 0.0);
    }
  }
  // block max of |q|
  __shared__ double red[32];
  for (int o = 16; o; o >>= 1) m = dmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    double mm = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) mm = dmaxf(mm, red[w]);
    P.qn[b] = mm; P.idx[b] = (int)o; P.done[b] = 0;
    for (int k = 0; k < 4; k++) P.red[b * 4 + k] = 0ULL;
  }
}

// ---- certificate partials: ndy = max|dy|, supp = sum hi*max(dy,0) + lo*min(dy,0); also writes DY (general cols only) for A' dy
struct CertParams {
  const double *YO, *YP, *QB, *lo, *hi;
  double* DY;       // [rows][NTp]: box cols 0, general cols dy
  double* cert;     // [rows][3]: ndy, supp, atdy
  int rows, NTp, nz, nt;
};
__global__ void stream_cert_kernel(const CertParams P) {
  const int b = blockIdx.x;
  double ndy = 0.0, supp = 0.0;
  for (int n = threadIdx.x; n < P.NTp; n += blockDim.x) {
    const size_t off = (size_t)b * P.NTp + n;
    double dy = 0.0;
    if (n < P.nt) {
      dy = P.YO[off] - P.YP[off];
      const double o = (n < P.nz) ? 0.0 : P.QB[off];
      ndy = dmaxf(ndy, fabs(dy));
      supp += (P.hi[n] + o) * dmaxf(dy, 0.0) + (P.lo[n] + o) * (dy < 0.0 ? dy : 0.0);
    }
    P.DY[off] = (n >= P.nz && n < P.nt) ? dy : 0.0;
  }
  __shared__ double s1[32], s2[32];
  for (int o = 16; o; o >>= 1) { ndy = dmaxf(ndy, __shfl_xor_sync(0xffffffffu, ndy, o)); supp += __shfl_xor_sync(0xffffffffu, supp, o); }
  if ((threadIdx.x & 31) == 0) { s1[threadIdx.x >> 5] = ndy; s2[threadIdx.x >> 5] = supp; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, c = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) { a = dmaxf(a, s1[w]); c += s2[w]; }
    P.cert[(size_t)b * 3 + 0] = a; P.cert[(size_t)b * 3 + 1] = c;
  }
}
// atdy = max over box cols |(G' dy_g)[n] + dy_box[n]|
__global__ void stream_cert2_kernel(const double* GtDy, const double* YO, const double* YP, double* cert, int NTp, int nz) {
  const int b = blockIdx.x;
  double m = 0.0;
  for (int n = threadIdx.x; n < nz; n += blockDim.x) {
    const size_t off = (size_t)b * NTp + n;
    m = dmaxf(m, fabs(GtDy[off] + (YO[off] - YP[off])));
  }
  __shared__ double s1[32];
  for (int o = 16; o; o >>= 1) m = dmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) s1[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) a = dmaxf(a, s1[w]);
    cert[(size_t)b * 3 + 2] = a;
  }
}

// ---- decide: per active row, OSQP criteria; finished rows are written out by stream_output_kernel --------------------
struct DecideParams {
  unsigned long long* red; const double* qn; const double* cert;   // cert may be null (mg == 0)
  int* idx; int* done; int* newly;
  int32_t* status; int32_t* iters; double* pres; double* dres;
  int rows, it, max_iter, iters_add;
  double eps_abs, eps_rel, eps_pinf;
  int* count;   // [0] += rows still active
};
__global__ void stream_decide_kernel(const DecideParams P) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.rows) return;
  P.newly[b] = 0;
  if (P.done[b]) return;
  unsigned long long* r4 = P.red + (size_t)b * 4;
  const double rp = __longlong_as_double((long long)r4[0]), rd = __longlong_as_double((long long)r4[1]);
  const double nA = __longlong_as_double((long long)r4[2]), nD = __longlong_as_double((long long)r4[3]);
  r4[0] = r4[1] = r4[2] = r4[3] = 0ULL;
  const bool conv = (rp <= P.eps_abs + P.eps_rel * nA) && (rd <= P.eps_abs + P.eps_rel * dmaxf(nD, P.qn[b]));
  bool pinf = false;
  if (P.cert != nullptr && !conv) {
    const double ndy = P.cert[(size_t)b * 3], supp = P.cert[(size_t)b * 3 + 1], atdy = P.cert[(size_t)b * 3 + 2];
    pinf = (ndy > P.eps_pinf) && (supp < -P.eps_pinf * ndy) && (atdy <= P.eps_pinf * ndy);
  }
  if (conv || pinf || P.it >= P.max_iter) {
    const int o = P.idx[b];
    P.status[o] = conv ? 1 : (pinf ? -3 : -2);
    P.iters[o] = P.it + P.iters_add; P.pres[o] = rp; P.dres[o] = rd;
    P.done[b] = 1; P.newly[b] = 1;
  } else {
    atomicAdd(P.count, 1);
  }
}
__global__ void stream_output_kernel(const int* newly, const int* idx, const double* XT, const double* YO, double* v_out, double* y_out,
                                     int NTp, int nz, int nt) {
  const int b = blockIdx.x;
  if (!newly[b]) return;
  const size_t o = (size_t)idx[b];
  for (int n = threadIdx.x; n < nz; n += blockDim.x) v_out[o * nz + n] = XT[(size_t)b * NTp + n];
  if (y_out != nullptr)
    for (int n = threadIdx.x; n < nt; n += blockDim.x) y_out[o * nt + n] = YO[(size_t)b * NTp + n];
}
// unordered compaction of the still-active rows into the alternate buffers
struct CompactParams {
  const int* done; const int* idx; int* idx2; int* done2;
  const double *Cst, *QB, *X, *R, *qn;
  double *Cst2, *QB2, *X2, *R2, *qn2;
  unsigned long long* red2;
  int NTp; int* counter; int sig;
};
__global__ void stream_compact_kernel(const CompactParams P) {
  const int b = blockIdx.x;
  if (P.done[b]) return;
  __shared__ int pos;
  if (threadIdx.x == 0) {
    pos = atomicAdd(P.counter, 1);
    P.idx2[pos] = P.idx[b]; P.done2[pos] = 0; P.qn2[pos] = P.qn[b];
    for (int k = 0; k < 4; k++) P.red2[(size_t)pos * 4 + k] = 0ULL;
  }
  __syncthreads();
  const size_t src = (size_t)b * P.NTp, dst = (size_t)pos * P.NTp;
  for (int n = threadIdx.x * 2; n < P.NTp; n += blockDim.x * 2) {
    *reinterpret_cast<double2*>(P.Cst2 + dst + n) = *reinterpret_cast<const double2*>(P.Cst + src + n);
    *reinterpret_cast<double2*>(P.QB2 + dst + n) = *reinterpret_cast<const double2*>(P.QB + src + n);
    *reinterpret_cast<double2*>(P.R2 + dst + n) = *reinterpret_cast<const double2*>(P.R + src + n);
    if (P.sig) *reinterpret_cast<double2*>(P.X2 + dst + n) = *reinterpret_cast<const double2*>(P.X + src + n);
  }
}

cudaError_t dev_alloc(double** p, size_t n) { return cudaMalloc(p, std::max<size_t>(n, 1) * sizeof(double)); }

}  // namespace

int stream_padded(int nt) { return ((nt + BN - 1) / BN) * BN; }

cudaError_t stream_upload(const Design& D, StreamConsts& sc, std::string& err) {
  const int NTp = stream_padded(D.nt), nt = D.nt, np = D.np;
  sc.NTp = NTp;
  std::vector<double> T((size_t)NTp * NTp, 0.0), C((size_t)NTp * NTp, 0.0), Lt((size_t)np * NTp, 0.0), Lv(D.Lv.a.empty() ? 0 : (size_t)np * NTp, 0.0), lo(NTp, 0.0), hi(NTp, 0.0), rho(NTp, 1.0),
      rinv(NTp, 1.0);
  for (int j = 0; j < nt; j++)
    for (int i = 0; i < nt; i++) { T[(size_t)i * NTp + j] = D.T(i, j); C[(size_t)i * NTp + j] = D.C(i, j); }
  for (int j = 0; j < np; j++) {
    for (int i = 0; i < D.nz; i++) Lt[(size_t)j * NTp + i] = D.Lq(i, j);
    for (int i = 0; i < D.mg; i++) Lt[(size_t)j * NTp + D.nz + i] = D.Lb(i, j);
    if (!Lv.empty()) for (int i = 0; i < D.nz; i++) Lv[(size_t)j * NTp + i] = D.Lv(i, j);
  }
  for (int i = 0; i < nt; i++) { lo[i] = D.lo[i]; hi[i] = D.hi[i]; rho[i] = D.rho_vec[i]; rinv[i] = 1.0 / D.rho_vec[i]; }
  struct { double** p; const std::vector<double>* v; } ups[] = {{&sc.T, &T}, {&sc.C, &C}, {&sc.Lt, &Lt}, {&sc.Lv, &Lv}, {&sc.lo, &lo}, {&sc.hi, &hi}, {&sc.rho, &rho}, {&sc.rinv, &rinv}};
  for (auto& u : ups) {
    if (u.v->empty()) { *u.p = nullptr; continue; }
    cudaError_t e = dev_alloc(u.p, u.v->size());
    if (e != cudaSuccess) { err = "alloc constants"; return e; }
    e = cudaMemcpy(*u.p, u.v->data(), u.v->size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { err = "copy constants"; return e; }
  }
  cudaError_t e = cudaFuncSetAttribute(stream_iter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
  if (e != cudaSuccess) { err = "smem attr"; return e; }
  e = cudaFuncSetAttribute(stream_check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
  if (e != cudaSuccess) { err = "smem attr"; return e; }
  return cudaSuccess;
}

static cudaError_t ensure_work(StreamWork& sw, long long rows, int NTp, int check_every) {
  if (rows <= sw.cap && sw.NTp == NTp && 1 + (size_t)check_every * ((rows + BM - 1) / BM) <= sw.sched_cap) return cudaSuccess;
  double** arrs[] = {&sw.X, &sw.Q, &sw.Z, &sw.YS, &sw.R0, &sw.R1, &sw.DY, &sw.X2, &sw.Q2, &sw.Z2, &sw.XT, &sw.YO, &sw.YP};
  for (auto a : arrs) { if (*a) cudaFree(*a); *a = nullptr; }
  for (auto a : arrs) { cudaError_t e = dev_alloc(a, (size_t)rows * NTp); if (e != cudaSuccess) return e; }
  double** small[] = {&sw.qn, &sw.qn2, &sw.cert};
  for (auto a : small) { if (*a) cudaFree(*a); *a = nullptr; cudaError_t e = dev_alloc(a, (size_t)rows * 3); if (e != cudaSuccess) return e; }
  int** ints[] = {&sw.idx, &sw.idx2, &sw.done, &sw.done2, &sw.newly};
  for (auto a : ints) { if (*a) cudaFree(*a); *a = nullptr; cudaError_t e = cudaMalloc(a, std::max<size_t>(rows, 1) * sizeof(int)); if (e != cudaSuccess) return e; }
  if (sw.red) cudaFree(sw.red); if (sw.red2) cudaFree(sw.red2);
  cudaError_t e = cudaMalloc(&sw.red, std::max<size_t>(rows, 1) * 4 * sizeof(unsigned long long)); if (e != cudaSuccess) return e;
  e = cudaMalloc(&sw.red2, std::max<size_t>(rows, 1) * 4 * sizeof(unsigned long long)); if (e != cudaSuccess) return e;
  if (!sw.count) { e = cudaMalloc(&sw.count, 2 * sizeof(int)); if (e != cudaSuccess) return e; }
  if (sw.sched) cudaFree(sw.sched);
  sw.sched = nullptr; sw.sched_cap = 1 + (size_t)check_every * ((rows + BM - 1) / BM);
  e = cudaMalloc(&sw.sched, sw.sched_cap * sizeof(int)); if (e != cudaSuccess) return e;
  {
    int occ = 0, dev = 0, sms = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, stream_iter_kernel, THREADS, SMEM_BYTES); if (e != cudaSuccess) return e;
    cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    sw.resident_ctas = std::max(1, occ) * sms;      // the ticket scheme needs every CTA of the grid resident
  }
  if (!sw.h_count) { e = cudaMallocHost(&sw.h_count, 2 * sizeof(int)); if (e != cudaSuccess) return e; }
  sw.cap = rows; sw.NTp = NTp;
  return cudaSuccess;
}

cudaError_t stream_solve(const Design& D, const mpcb_settings& st, const StreamConsts& sc, StreamWork& sw, const StreamBatch& B, int sm_count,
                         cudaStream_t stream, int* launches, std::string& err) {
  (void)sm_count;
  const int NTp = sc.NTp, nz = D.nz, nt = D.nt, mg = D.mg;
  if (B.batch > 0x7fffffffLL / 4) { err = "batch too large"; return cudaErrorInvalidValue; }
  cudaError_t e = ensure_work(sw, B.batch, NTp, st.check_every);
  if (e != cudaSuccess) { err = "workspace allocation"; return e; }
  int nl = 0;
  int rows = (int)B.batch;
  // state buffers (sw.Z holds c, sw.Q holds q/b, sw.X holds x); alternates for compaction
  double *Cst = sw.Z, *QB = sw.Q, *X = sw.X, *Cst2 = sw.Z2, *QB2 = sw.Q2, *X2 = sw.X2, *Rin = sw.R0, *Rout = sw.R1;
  double *qn = sw.qn, *qn2 = sw.qn2;
  int *idx = sw.idx, *idx2 = sw.idx2, *done = sw.done, *done2 = sw.done2;
  unsigned long long *red = sw.red, *red2 = sw.red2;
  {
    InitParams P;
    P.Lt = sc.Lt; P.x0 = B.x0; P.xref = B.xref; P.uref = B.uref; P.xref_bc = B.xref_bc; P.uref_bc = B.uref_bc; P.nx = D.nx; P.nu = D.nu;
    P.np = D.np; P.NTp = NTp; P.nz = nz; P.nt = nt; P.rows = rows; P.Cst = Cst; P.QB = QB; P.X = X; P.R = Rin; P.qn = qn; P.idx = idx; P.done = done;
    P.red = red;
    P.warm_v = B.warm_v; P.warm_y = B.warm_y; P.C = sc.C; P.rho = sc.rho; P.alpha = st.alpha; P.sigma = st.sigma; P.remap = B.remap;
    P.Lv = (st.cold_init && sc.Lv) ? sc.Lv : nullptr; P.lo = sc.lo; P.hi = sc.hi;
    stream_init_kernel<<<rows, 128, (D.np + 2 * (size_t)nz) * sizeof(double), stream>>>(P); nl++;
  }
  const int max_iter = ((st.max_iter + st.check_every - 1) / st.check_every) * st.check_every;
  const bool sig = st.sigma != 0.0;
  int it = 0, n_done_rows = 0;   // rows finished but not yet compacted away
  while (rows > 0 && it < max_iter) {
    {
      IterParams P;
      P.R0 = Rin; P.R1 = Rout; P.iters = st.check_every; P.sched = sw.sched;
      P.nrb = (rows + BM - 1) / BM; P.ncb = NTp / BN;
      P.T = sc.T; P.Cst = Cst; P.QB = QB; P.X = X; P.lo = sc.lo; P.hi = sc.hi; P.rho = sc.rho;
      P.XT = sw.XT; P.YO = sw.YO; P.YP = sw.YP; P.red = red; P.rows = rows; P.NTp = NTp; P.nz = nz; P.nt = nt; P.mg = mg;
      P.alpha = st.alpha; P.sigma = st.sigma;
      e = cudaMemsetAsync(sw.sched, 0, sizeof(int) * (1 + (size_t)st.check_every * P.nrb), stream);
      if (e != cudaSuccess) { err = "memset"; return e; }
      const int tiles = P.nrb * P.ncb;
      const int grid = std::min(tiles * st.check_every, sw.resident_ctas);
      stream_iter_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(P); nl++;
      if (st.check_every & 1) std::swap(Rin, Rout);      // the operand of the next period is where the last iteration wrote
    }
    it += st.check_every;
    if (mg > 0) {
      // check operand [x~; y_g]: box cols from XT, general cols from YO -> assemble into DY buffer, then one pass with C
      CertParams cp;   // reuse the kernel below for dy as well (needs YP of the previous iteration; check_every >= 2)
      cp.YO = sw.YO; cp.YP = sw.YP; cp.QB = QB; cp.lo = sc.lo; cp.hi = sc.hi; cp.DY = sw.DY; cp.cert = sw.cert; cp.rows = rows; cp.NTp = NTp; cp.nz = nz; cp.nt = nt;
      if (st.check_every >= 2) { stream_cert_kernel<<<rows, 128, 0, stream>>>(cp); nl++; }
      // both check products are needed on the box columns only (G' dy_g and Pc x~ + G' y_g have nz rows); the certificate operand DY is zero on the
      // box columns, so its product starts at the first k-tile that holds a general column
      dim3 grid((nz + BN - 1) / BN, (rows + BM - 1) / BM);
      if (st.check_every >= 2) {   // G' dy_g into Rout (scratch: it is fully rewritten by the next iteration)
        CheckParams c2; c2.In = sw.DY; c2.C = sc.C; c2.QB = QB; c2.YO = sw.YO; c2.red = red; c2.Out = Rout; c2.rows = rows; c2.NTp = NTp; c2.nz = nz; c2.mode = 1;
        c2.kt0 = nz / BK;
        c2.done = done; c2.cert = sw.cert; c2.eps_abs = st.eps_abs; c2.eps_rel = st.eps_rel; c2.eps_pinf = st.eps_prim_inf; c2.want_all = 0;
        stream_check_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(c2); nl++;
        stream_cert2_kernel<<<rows, 128, 0, stream>>>(Rout, sw.YO, sw.YP, sw.cert, NTp, nz); nl++;
      }
      // operand for the residual pass: XT on box cols, YO on general cols (assembled in DY)
      // (a tiny fused copy: reuse compact-style loop through cudaMemcpy2DAsync)
      e = cudaMemcpy2DAsync(sw.DY, NTp * sizeof(double), sw.XT, NTp * sizeof(double), nz * sizeof(double), rows, cudaMemcpyDeviceToDevice, stream);
      if (e != cudaSuccess) { err = "memcpy2d"; return e; }
      e = cudaMemcpy2DAsync(sw.DY + nz, NTp * sizeof(double), sw.YO + nz, NTp * sizeof(double), (NTp - nz) * sizeof(double), rows, cudaMemcpyDeviceToDevice, stream);
      if (e != cudaSuccess) { err = "memcpy2d"; return e; }
      CheckParams c1; c1.In = sw.DY; c1.C = sc.C; c1.QB = QB; c1.YO = sw.YO; c1.red = red; c1.Out = nullptr; c1.rows = rows; c1.NTp = NTp; c1.nz = nz; c1.mode = 0; c1.kt0 = 0;
      c1.done = done; c1.cert = (st.check_every >= 2) ? sw.cert : nullptr; c1.eps_abs = st.eps_abs; c1.eps_rel = st.eps_rel; c1.eps_pinf = st.eps_prim_inf; c1.want_all = it >= max_iter ? 1 : 0;
      stream_check_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(c1); nl++;
    }
    e = cudaMemsetAsync(sw.count, 0, 2 * sizeof(int), stream);
    if (e != cudaSuccess) { err = "memset"; return e; }
    DecideParams dp;
    dp.red = red; dp.qn = qn; dp.cert = (mg > 0 && st.check_every >= 2) ? sw.cert : nullptr; dp.idx = idx; dp.done = done; dp.newly = sw.newly;
    dp.status = B.status; dp.iters = B.iters; dp.pres = B.pres; dp.dres = B.dres; dp.rows = rows; dp.it = it; dp.max_iter = max_iter; dp.iters_add = B.iters_add;
    dp.eps_abs = st.eps_abs; dp.eps_rel = st.eps_rel; dp.eps_pinf = st.eps_prim_inf; dp.count = sw.count;
    stream_decide_kernel<<<(rows + 127) / 128, 128, 0, stream>>>(dp); nl++;
    stream_output_kernel<<<rows, 128, 0, stream>>>(sw.newly, idx, sw.XT, sw.YO, B.v_out, B.y_out, NTp, nz, nt); nl++;
    e = cudaMemcpyAsync(sw.h_count, sw.count, sizeof(int), cudaMemcpyDeviceToHost, stream);
    if (e != cudaSuccess) { err = "count readback"; return e; }
    e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) { err = "iteration kernels"; return e; }
    const int active = sw.h_count[0];
    n_done_rows = rows - active;
    if (active == 0) break;
    if (n_done_rows * 8 >= rows) {   // compact when >= 1/8 of the rows are finished
      CompactParams cp;
      cp.done = done; cp.idx = idx; cp.idx2 = idx2; cp.done2 = done2; cp.Cst = Cst; cp.QB = QB; cp.X = X; cp.R = Rin; cp.qn = qn;
      cp.Cst2 = Cst2; cp.QB2 = QB2; cp.X2 = X2; cp.R2 = Rout; cp.qn2 = qn2; cp.red2 = red2; cp.NTp = NTp; cp.counter = sw.count + 1; cp.sig = sig ? 1 : 0;
      stream_compact_kernel<<<rows, 128, 0, stream>>>(cp); nl++;
      std::swap(Cst, Cst2); std::swap(QB, QB2); std::swap(X, X2); std::swap(Rin, Rout); std::swap(qn, qn2); std::swap(idx, idx2); std::swap(done, done2);
      std::swap(red, red2);
      rows = active;
    }
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) { err = "kernel launch"; return e; }
  if (launches) *launches = nl;
  return cudaSuccess;
}

void stream_release(StreamConsts& sc, StreamWork& sw) {
  double** cs[] = {&sc.T, &sc.C, &sc.Lt, &sc.Lv, &sc.lo, &sc.hi, &sc.rho, &sc.rinv};
  for (auto p : cs) { if (*p) cudaFree(*p); *p = nullptr; }
  double** ws[] = {&sw.X, &sw.Q, &sw.Z, &sw.YS, &sw.R0, &sw.R1, &sw.DY, &sw.X2, &sw.Q2, &sw.Z2, &sw.XT, &sw.YO, &sw.YP, &sw.qn, &sw.qn2, &sw.cert};
  for (auto p : ws) { if (*p) cudaFree(*p); *p = nullptr; }
  int** is[] = {&sw.idx, &sw.idx2, &sw.done, &sw.done2, &sw.newly, &sw.count, &sw.sched};
  for (auto p : is) { if (*p) cudaFree(*p); *p = nullptr; }
  if (sw.red) cudaFree(sw.red); if (sw.red2) cudaFree(sw.red2); sw.red = sw.red2 = nullptr;
  if (sw.h_count) cudaFreeHost(sw.h_count); sw.h_count = nullptr;
  sw.cap = 0;
}

}  // namespace mpcb
