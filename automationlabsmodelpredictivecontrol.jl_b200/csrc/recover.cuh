// Rebuild the reference's result matrices from the optimal absolute inputs (sm_100a, HBM-bound).
//
// `calculate!` copies u, e_u, x, e_x out of the solver (/root/reference/src/main/computation_mpc.jl:50-53).  The
// condensed solve only carries v = vec(u); this kernel rolls the reference's deviation dynamics
// e_{k+1} = A e_k + B (u_k - u_ref)  (linear/mpc_modeler_implementation_linear.jl:59) once per problem, writes the
// four result matrices in the reference's per-problem column-major layout and evaluates the reference's cost J
// (src/sub/design_mpc.jl:449-456) with all constants.  One thread per problem; the running deviation lives in
// shared memory as [nx][blockDim] columns (bank-conflict free), every store is a full 8*nx / 8*nu byte run.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mpcb {

struct RecoverParams {
  const double* A;  // nx*nx column-major
  const double* B;  // nx*nu
  const double* Q;  // nx*nx
  const double* R;  // nu*nu
  const double* S;  // nu*nu
  const double* Pt; // nx*nx terminal
  int nx, nu, H, use_R, use_S;
  long long batch;
  const double* x0;
  const double* xref;
  const double* uref;
  int xref_bc, uref_bc;
  const double* v;  // [batch][nu*H]
  double* u;        // [batch][H][nu]      (nu x H column-major per problem)
  double* e_u;
  double* x;        // [batch][H+1][nx]
  double* e_x;
  double* u0;       // [batch][nu]
  double* objective;
};

constexpr int RECOVER_THREADS = 128;

__host__ __device__ inline size_t recover_smem_bytes(int nx, int nu, int threads = RECOVER_THREADS) {
  return sizeof(double) * ((size_t)(2 * nx + 2 * nu) * threads + 3 * nx * nx + nx * nu + 2 * nu * nu);
}
// CTA width of the generic kernel: the per-thread deviation columns live in shared memory, so wide systems run
// narrower CTAs (nx = 64, nu = 16: 64 threads, 192 KB).  Returns 0 when even one warp does not fit.
inline int recover_threads_for(int nx, int nu) {
  for (int t = RECOVER_THREADS; t >= 32; t >>= 1)
    if (recover_smem_bytes(nx, nu, t) <= 200 * 1024) return t;
  return 0;
}

__global__ void __launch_bounds__(RECOVER_THREADS) recover_kernel(const RecoverParams P) {
  extern __shared__ __align__(16) double sm[];
  const int nx = P.nx, nu = P.nu, H = P.H, tid = threadIdx.x, TT = blockDim.x;
  double* se = sm;                                   // [nx][T] current deviation
  double* sn = se + nx * TT;            // [nx][T] next deviation
  double* su = sn + nx * TT;            // [nu][T] current input deviation
  double* sp = su + nu * TT;            // [nu][T] previous absolute input (S term)
  double* sA = sp + nu * TT;
  double* sB = sA + nx * nx;
  double* sQ = sB + nx * nu;
  double* sPt = sQ + nx * nx;
  double* sR = sPt + nx * nx;
  double* sS = sR + nu * nu;
  for (int i = tid; i < nx * nx; i += TT) { sA[i] = P.A[i]; sQ[i] = P.Q[i]; sPt[i] = P.Pt[i]; }
  for (int i = tid; i < nx * nu; i += TT) sB[i] = P.B[i];
  for (int i = tid; i < nu * nu; i += TT) { sR[i] = P.R[i]; sS[i] = P.S ? P.S[i] : 0.0; }
  __syncthreads();
  const long long p = (long long)blockIdx.x * TT + tid;
  if (p >= P.batch) return;
  const double* x0 = P.x0 + p * nx;
  const double* xr = P.xref + (P.xref_bc ? 0 : p) * nx;
  const double* ur = P.uref + (P.uref_bc ? 0 : p) * nu;
  const double* v = P.v + p * (long long)nu * H;
  for (int i = 0; i < nx; i++) se[i * TT + tid] = x0[i] - xr[i];
  double J = 0.0;
  for (int k = 0; k <= H; k++) {
    // write x_k, e_x_k and accumulate e' W e
    const double* W = (k == H) ? sPt : sQ;
    double quad = 0.0;
    for (int i = 0; i < nx; i++) {
      const double ei = se[i * TT + tid];
      if (P.e_x) P.e_x[(p * (H + 1) + k) * nx + i] = ei;
      if (P.x) P.x[(p * (H + 1) + k) * nx + i] = ei + xr[i];
      double s = 0.0;
      for (int j = 0; j < nx; j++) s = fma(W[j * nx + i], se[j * TT + tid], s);
      quad = fma(ei, s, quad);
    }
    J += quad;
    if (k == H) break;
    for (int i = 0; i < nu; i++) {
      const double ui = v[k * nu + i];
      const double eu = ui - ur[i];
      su[i * TT + tid] = eu;
      if (P.u) P.u[(p * H + k) * nu + i] = ui;
      if (P.e_u) P.e_u[(p * H + k) * nu + i] = eu;
      if (k == 0 && P.u0) P.u0[p * nu + i] = ui;
    }
    if (P.use_R) {
      double quadr = 0.0;
      for (int i = 0; i < nu; i++) {
        double s = 0.0;
        for (int j = 0; j < nu; j++) s = fma(sR[j * nu + i], su[j * TT + tid], s);
        quadr = fma(su[i * TT + tid], s, quadr);
      }
      J += quadr;
      if (P.use_S) {
        if (k > 0) {  // delta_u_{k-1} = u_{k-1} - u_k  (design_mpc.jl:429-432)
          double quads = 0.0;
          for (int i = 0; i < nu; i++) {
            double s = 0.0;
            for (int j = 0; j < nu; j++) s = fma(sS[j * nu + i], sp[j * TT + tid] - v[k * nu + j], s);
            quads = fma(sp[i * TT + tid] - v[k * nu + i], s, quads);
          }
          J += quads;
        }
        for (int i = 0; i < nu; i++) sp[i * TT + tid] = v[k * nu + i];
      }
    }
    for (int i = 0; i < nx; i++) {
      double s = 0.0;
      for (int j = 0; j < nx; j++) s = fma(sA[j * nx + i], se[j * TT + tid], s);
      for (int j = 0; j < nu; j++) s = fma(sB[j * nx + i], su[j * TT + tid], s);
      sn[i * TT + tid] = s;
    }
    double* tmp = se; se = sn; sn = tmp;
  }
  if (P.objective) P.objective[p] = J;
}

// Register-resident specialisation for small systems (the quadruple tank is NX=4, NU=2): the deviation state, the
// references and the weights' rows live in registers, every result column leaves as 16-byte stores that fill whole
// 32-byte sectors, and the per-problem input row is fetched with 16-byte loads.  The rollout is a dependent chain, so the
// inputs are prefetched RCH steps at a time, one chunk ahead of the chunk being rolled: the loads of chunk c+1 are in
// flight while chunk c computes and stores (ncu on the first version: 24 long-scoreboard stall cycles per issue).
template <int NX, int NU>
__global__ void __launch_bounds__(RECOVER_THREADS) recover_small_kernel(const RecoverParams P) {
  __shared__ double sA[NX * NX], sB[NX * NU], sQ[NX * NX], sPt[NX * NX], sR[NU * NU], sS[NU * NU];
  const int tid = threadIdx.x, H = P.H;
  for (int i = tid; i < NX * NX; i += RECOVER_THREADS) { sA[i] = P.A[i]; sQ[i] = P.Q[i]; sPt[i] = P.Pt[i]; }
  for (int i = tid; i < NX * NU; i += RECOVER_THREADS) sB[i] = P.B[i];
  for (int i = tid; i < NU * NU; i += RECOVER_THREADS) { sR[i] = P.R[i]; sS[i] = P.S ? P.S[i] : 0.0; }
  __syncthreads();
  const long long p = (long long)blockIdx.x * RECOVER_THREADS + tid;
  if (p >= P.batch) return;
  constexpr int NM = NX > NU ? NX : NU;
  constexpr int RCH = 4;
  double e[NX], xr[NX], ur[NU], up[NU];
#pragma unroll
  for (int i = 0; i < NX; i++) { xr[i] = P.xref[(P.xref_bc ? 0 : p) * NX + i]; e[i] = P.x0[p * NX + i] - xr[i]; }
#pragma unroll
  for (int i = 0; i < NU; i++) { ur[i] = P.uref[(P.uref_bc ? 0 : p) * NU + i]; up[i] = 0.0; }
  const double* __restrict__ v = P.v + p * (long long)NU * H;
  double* __restrict__ o_u = P.u ? P.u + p * (long long)H * NU : nullptr;
  double* __restrict__ o_eu = P.e_u ? P.e_u + p * (long long)H * NU : nullptr;
  double* __restrict__ o_x = P.x ? P.x + p * (long long)(H + 1) * NX : nullptr;
  double* __restrict__ o_ex = P.e_x ? P.e_x + p * (long long)(H + 1) * NX : nullptr;
  const bool vec_in = (NU & 1) == 0 && ((reinterpret_cast<uintptr_t>(v) & 15) == 0);
  auto load_chunk = [&](int c, double (&dst)[RCH * NU]) {
    const int k0 = c * RCH;
    if (vec_in && k0 + RCH <= H) {
#pragma unroll
      for (int i = 0; i < RCH * NU; i += 2) { const double2 t = *reinterpret_cast<const double2*>(v + k0 * NU + i); dst[i] = t.x; dst[i + 1] = t.y; }
    } else {
#pragma unroll
      for (int i = 0; i < RCH * NU; i++) dst[i] = (k0 * NU + i < H * NU) ? v[k0 * NU + i] : 0.0;
    }
  };
  auto store_run = [](double* dst, const double (&val)[NM], int n) {
    if ((n & 1) == 0 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
      for (int i = 0; i < NM; i += 2)
        if (i < n) *reinterpret_cast<double2*>(dst + i) = make_double2(val[i], val[i + 1]);
    } else {
#pragma unroll
      for (int i = 0; i < NM; i++)
        if (i < n) dst[i] = val[i];
    }
  };
  double J = 0.0;
  // writes x_k, e_x_k and returns e' W e
  auto emit_state = [&](int k, const double* W) {
    double tmp[NM];
    if (o_ex) {
#pragma unroll
      for (int i = 0; i < NX; i++) tmp[i] = e[i];
      store_run(o_ex + k * NX, tmp, NX);
    }
    if (o_x) {
#pragma unroll
      for (int i = 0; i < NX; i++) tmp[i] = e[i] + xr[i];
      store_run(o_x + k * NX, tmp, NX);
    }
    double quad = 0.0;
#pragma unroll
    for (int i = 0; i < NX; i++) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < NX; j++) s = fma(W[j * NX + i], e[j], s);
      quad = fma(e[i], s, quad);
    }
    return quad;
  };
  double cur[RCH * NU], nxt[RCH * NU];
  load_chunk(0, cur);
  const int nchunks = (H + RCH - 1) / RCH;
  for (int c = 0; c < nchunks; c++) {
    if (c + 1 < nchunks) load_chunk(c + 1, nxt);
#pragma unroll
    for (int q = 0; q < RCH; q++) {
      const int k = c * RCH + q;
      if (k < H) {
        J += emit_state(k, sQ);
        double uk[NM], eu[NM];
#pragma unroll
        for (int i = 0; i < NU; i++) { uk[i] = cur[q * NU + i]; eu[i] = uk[i] - ur[i]; }
        if (o_u) store_run(o_u + k * NU, uk, NU);
        if (o_eu) store_run(o_eu + k * NU, eu, NU);
        if (k == 0 && P.u0) store_run(P.u0 + p * NU, uk, NU);
        if (P.use_R) {
          double quadr = 0.0;
#pragma unroll
          for (int i = 0; i < NU; i++) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < NU; j++) s = fma(sR[j * NU + i], eu[j], s);
            quadr = fma(eu[i], s, quadr);
          }
          J += quadr;
          if (P.use_S) {
            if (k > 0) {        // delta_u_{k-1} = u_{k-1} - u_k  (design_mpc.jl:429-432)
              double quads = 0.0;
#pragma unroll
              for (int i = 0; i < NU; i++) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j < NU; j++) s = fma(sS[j * NU + i], up[j] - uk[j], s);
                quads = fma(up[i] - uk[i], s, quads);
              }
              J += quads;
            }
#pragma unroll
            for (int i = 0; i < NU; i++) up[i] = uk[i];
          }
        }
        double en[NX];
#pragma unroll
        for (int i = 0; i < NX; i++) {
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < NX; j++) s = fma(sA[j * NX + i], e[j], s);
#pragma unroll
          for (int j = 0; j < NU; j++) s = fma(sB[j * NX + i], eu[j], s);
          en[i] = s;
        }
#pragma unroll
        for (int i = 0; i < NX; i++) e[i] = en[i];
      }
    }
#pragma unroll
    for (int i = 0; i < RCH * NU; i++) cur[i] = nxt[i];
  }
  J += emit_state(H, sPt);
  if (P.objective) P.objective[p] = J;
}

// returns false when no specialisation exists (caller falls back to recover_kernel)
inline bool launch_recover_small(const RecoverParams& R, cudaStream_t st) {
  const unsigned grid = (unsigned)((R.batch + RECOVER_THREADS - 1) / RECOVER_THREADS);
#define MPCB_RS(NX_, NU_) \
  if (R.nx == NX_ && R.nu == NU_) { recover_small_kernel<NX_, NU_><<<grid, RECOVER_THREADS, 0, st>>>(R); return true; }
  MPCB_RS(2, 1) MPCB_RS(2, 2) MPCB_RS(3, 1) MPCB_RS(3, 2) MPCB_RS(4, 1) MPCB_RS(4, 2) MPCB_RS(4, 4) MPCB_RS(6, 2) MPCB_RS(6, 3)
  MPCB_RS(8, 2) MPCB_RS(8, 4)
#undef MPCB_RS
  return false;
}

}  // namespace mpcb
