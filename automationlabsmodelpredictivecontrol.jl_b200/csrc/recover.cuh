// Rebuild the reference's result matrices from the optimal absolute inputs (sm_100a, HBM-bound).
//
// `calculate!` copies u, e_u, x, e_x out of the solver (/root/reference/src/main/computation_mpc.jl:50-53).  The
// condensed solve only carries v = vec(u); this kernel rolls the reference's deviation dynamics
// e_{k+1} = A e_k + B (u_k - u_ref)  (linear/mpc_modeler_implementation_linear.jl:59) once per problem, writes the
// four result matrices in the reference's per-problem column-major layout and evaluates the reference's cost J
// (src/sub/design_mpc.jl:449-456) with all constants.  One thread per problem; the running deviation lives in
// shared memory as [nx][blockDim] columns (bank-conflict free), every store is a full 8*nx / 8*nu byte run.
#pragma once
#include <cstdlib>
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

// Wide systems (nx >= 16; config 3 is nx = 64, nu = 16): the one-thread-per-problem kernel above walks 51 x (64 x 80) multiply-adds
// serially per problem with 64 threads per CTA -- 6.9 ms for 8 192 problems, a fifth of the whole config-3 solve (launch list,
// profiles/r02/launches_lti64_tuned_summary.txt).  Here a GROUP of TP = ceil32(nx) threads owns a problem: thread i keeps row i of the
// recursion (e_{k+1}[i] = A[i,:] e_k + B[i,:] du_k) and of the cost (e_k[i] (W e_k)[i]); A, B, Q, P sit in shared memory column-major,
// so consecutive threads read consecutive addresses and every global store is a coalesced run of nx (nu) doubles.  Persistent CTAs of
// 256 threads loop over problems, so the matrices are staged once per CTA.
constexpr int RECOVER_WIDE_THREADS = 256;
__host__ __device__ inline size_t recover_wide_smem_bytes(int nx, int nu) {
  const int tp = ((nx + 31) / 32) * 32, gpc = RECOVER_WIDE_THREADS / tp;
  return sizeof(double) * ((size_t)3 * nx * nx + (size_t)nx * nu + 2 * (size_t)nu * nu + (size_t)gpc * (2 * tp + 3 * nu + 8));
}
inline bool recover_wide_applies(int nx, int nu) { return nx >= 16 && nx <= 256 && nu <= nx && recover_wide_smem_bytes(nx, nu) <= 200 * 1024; }

// NXB > 0: row i of A lives in NXB registers of thread i (nx <= NXB) -- one broadcast shared-memory read of e[j] per multiply-add instead of
// two reads: the first version (both operands from shared memory, and the full W e product for the cost) was shared-memory bound at 3 ms
// for config 3.  QDIAG: Q is diagonal (the reference only ever builds Q = mpc_Q * I, design_mpc.jl:264-283): the stage cost needs no product.
template <int NXB, bool QDIAG>
__global__ void __launch_bounds__(RECOVER_WIDE_THREADS) recover_wide_kernel(const RecoverParams P) {
  extern __shared__ __align__(16) double sm[];
  const int nx = P.nx, nu = P.nu, H = P.H, tid = threadIdx.x;
  const int tp = ((nx + 31) / 32) * 32, gpc = RECOVER_WIDE_THREADS / tp;       // threads per problem, problems per CTA
  double* sA = sm; double* sQ = sA + nx * nx; double* sPt = sQ + nx * nx; double* sB = sPt + nx * nx; double* sR = sB + nx * nu; double* sS = sR + nu * nu;
  double* grp = sS + nu * nu + (size_t)(tid / tp) * (2 * tp + 3 * nu + 8);
  double* se = grp; double* sn = se + tp; double* su = sn + tp; double* sp = su + nu; double* sv = sp + nu; double* sred = sv + nu;   // sred: 8 warp partials
  for (int i = tid; i < nx * nx; i += RECOVER_WIDE_THREADS) { sA[i] = P.A[i]; sQ[i] = P.Q[i]; sPt[i] = P.Pt[i]; }
  for (int i = tid; i < nx * nu; i += RECOVER_WIDE_THREADS) sB[i] = P.B[i];
  for (int i = tid; i < nu * nu; i += RECOVER_WIDE_THREADS) { sR[i] = P.R[i]; sS[i] = P.S ? P.S[i] : 0.0; }
  __syncthreads();
  const int g = tid / tp, i = tid % tp;
  const bool row = i < nx, urow = i < nu, live = g < gpc;
  double areg[NXB > 0 ? NXB : 1];
  double qii = 0.0;
  if (NXB > 0) {
#pragma unroll
    for (int j = 0; j < NXB; j++) areg[j] = (row && j < nx) ? sA[j * nx + i] : 0.0;
  }
  if (QDIAG && row) qii = sQ[i * nx + i];
  const long long stride = (long long)gridDim.x * gpc;
  const long long rounds = (P.batch + stride - 1) / stride;
  for (long long r = 0; r < rounds; r++) {
    const long long p = r * stride + (long long)blockIdx.x * gpc + g;
    const bool act = live && p < P.batch;
    const long long pc = act ? p : 0;
    const double* xr = P.xref + (P.xref_bc ? 0 : pc) * nx;
    const double* ur = P.uref + (P.uref_bc ? 0 : pc) * nu;
    const double* v = P.v + pc * (long long)nu * H;
    double xri = 0.0, uri = 0.0;
    if (live) { if (row) { xri = xr[i]; se[i] = P.x0[pc * nx + i] - xri; } else { se[i] = 0.0; sn[i] = 0.0; } }      // rows nx .. tp-1 stay zero (register-blocked products run to NXB)
    if (live && urow) uri = ur[i];
    __syncthreads();
    double J = 0.0;
    for (int k = 0; k <= H; k++) {
      const double* W = (k == H) ? sPt : sQ;
      if (live && row) {
        const double ei = se[i];
        if (act) {
          if (P.e_x) P.e_x[(p * (H + 1) + k) * nx + i] = ei;
          if (P.x) P.x[(p * (H + 1) + k) * nx + i] = ei + xri;
        }
        if (QDIAG && k < H) J = fma(qii * ei, ei, J);
        else {
          double s = 0.0;
          for (int j = 0; j < nx; j++) s = fma(W[j * nx + i], se[j], s);
          J = fma(ei, s, J);
        }
      }
      if (k == H) break;
      if (live && urow) {
        const double ui = v[k * nu + i], eu = ui - uri;
        su[i] = eu; sv[i] = ui;
        if (act) {
          if (P.u) P.u[(p * H + k) * nu + i] = ui;
          if (P.e_u) P.e_u[(p * H + k) * nu + i] = eu;
          if (k == 0 && P.u0) P.u0[p * nu + i] = ui;
        }
      }
      __syncthreads();
      if (live && urow && P.use_R) {
        double s = 0.0;
        for (int j = 0; j < nu; j++) s = fma(sR[j * nu + i], su[j], s);
        J = fma(su[i], s, J);
        if (P.use_S && k > 0) {      // delta_u_{k-1} = u_{k-1} - u_k  (design_mpc.jl:429-432)
          double t = 0.0;
          for (int j = 0; j < nu; j++) t = fma(sS[j * nu + i], sp[j] - sv[j], t);
          J = fma(sp[i] - sv[i], t, J);
        }
      }
      if (live && row) {
        double s = 0.0, s1 = 0.0;
        if (NXB > 0) {
#pragma unroll
          for (int j = 0; j < NXB; j += 2) { s = fma(areg[j], se[j], s); s1 = fma(areg[j + 1], se[j + 1], s1); }      // (se is padded with zeros up to NXB)
          s += s1;
        } else {
          for (int j = 0; j < nx; j++) s = fma(sA[j * nx + i], se[j], s);
        }
        for (int j = 0; j < nu; j++) s = fma(sB[j * nx + i], su[j], s);
        sn[i] = s;
      }
      __syncthreads();
      if (live && urow) sp[i] = sv[i];
      if (live && row) se[i] = sn[i];
      __syncthreads();
    }
    // J: sum over the group's threads (warp shuffles, then the group's warps through shared memory)
    for (int o = 16; o; o >>= 1) J += __shfl_xor_sync(0xffffffffu, J, o);
    if (live && (i & 31) == 0) sred[i >> 5] = J;
    __syncthreads();
    if (act && i == 0 && P.objective) {
      double t = 0.0;
      for (int w = 0; w < tp / 32; w++) t += sred[w];
      P.objective[p] = t;
    }
    __syncthreads();
  }
}

inline const void* recover_wide_variant(int nx, bool qdiag) {
  if (nx <= 32) return qdiag ? (const void*)recover_wide_kernel<32, true> : (const void*)recover_wide_kernel<32, false>;
  if (nx <= 64) return qdiag ? (const void*)recover_wide_kernel<64, true> : (const void*)recover_wide_kernel<64, false>;
  return qdiag ? (const void*)recover_wide_kernel<0, true> : (const void*)recover_wide_kernel<0, false>;
}

// Specialisation for small systems (the quadruple tank is NX=4, NU=2): one lane per problem rolls the deviation dynamics
// in registers; everything that touches HBM goes through a per-warp shared-memory tile so that the global accesses are
// WARP-COOPERATIVE: a chunk of RCH steps of 32 problems is staged (inputs in, results out) and moved with 16-byte
// accesses in which consecutive lanes cover consecutive bytes of one problem's row (128-byte runs for x / e_x, 64-byte
// runs for u / e_u at NX=4, NU=2).  History (profiles/r01/recover_qt_h20_ncu_full*.txt): with per-lane 16-byte stores
// every store instruction touched 32 different 128-byte lines, the kernel sat on lg/mio-throttle and long-scoreboard
// stalls and reached 1.5 TB/s.
// RCH = steps per staged chunk, two instantiations.  Measured on B200 (QT, solve + recover): 65536 problems: 4 -> 0.577 ms (67 KB per
// CTA, 3 CTAs/SM: the 512 CTAs need 1.15 waves), 3 -> 0.572, 2 -> 0.560 (39 KB, 5 CTAs/SM, one wave), 1 -> 0.581; one problem (the
// closed-loop latency path): 4 -> 24 us, 2 -> 28 us (more passes over the horizon); 8192-problem chunks of the pipelined host entry:
// 2 -> 34 us.  So: 2 only when the 4-step version would not fit one wave (more than 3 CTAs x 148 SMs), 4 otherwise.
constexpr int RECOVER_RCH_BATCH = 2, RECOVER_RCH_FEW = 4;
template <int NX, int NU, int RCH>
__host__ __device__ constexpr int recover_small_warp_doubles() {
  return 32 * (2 * (RCH * NX + 2) + 3 * (RCH * NU + 2));
}

// moves a [32 problems][n doubles] tile (pitch `pitch`) between shared memory and rows of global memory that are
// `gstride` doubles apart; 16-byte accesses when the row geometry allows it
template <bool STORE>
__device__ __forceinline__ void warp_tile_copy(double* tile, int pitch, double* gbase, long long gstride, int n, int nvalid, int lane, bool vec_ok) {
  if (vec_ok) {
    const int cpr = n >> 1;                                   // 16-byte chunks per row
    for (int idx = lane; idx < nvalid * cpr; idx += 32) {
      const int r = idx / cpr, c = idx - r * cpr;
      double2* g = reinterpret_cast<double2*>(gbase + r * gstride) + c;
      double2* t = reinterpret_cast<double2*>(tile + r * pitch) + c;
      if (STORE) *g = *t; else *t = *g;
    }
  } else {
    for (int idx = lane; idx < nvalid * n; idx += 32) {
      const int r = idx / n, c = idx - r * n;
      if (STORE) gbase[r * gstride + c] = tile[r * pitch + c]; else tile[r * pitch + c] = gbase[r * gstride + c];
    }
  }
}

template <int NX, int NU, int RCH>
__global__ void __launch_bounds__(RECOVER_THREADS) recover_small_kernel(const RecoverParams P) {
  __shared__ double sA[NX * NX], sB[NX * NU], sQ[NX * NX], sPt[NX * NX], sR[NU * NU], sS[NU * NU];
  extern __shared__ __align__(16) double tiles[];
  constexpr int PX = RCH * NX + 2, PU = RCH * NU + 2;
  const int tid = threadIdx.x, H = P.H, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < NX * NX; i += RECOVER_THREADS) { sA[i] = P.A[i]; sQ[i] = P.Q[i]; sPt[i] = P.Pt[i]; }
  for (int i = tid; i < NX * NU; i += RECOVER_THREADS) sB[i] = P.B[i];
  for (int i = tid; i < NU * NU; i += RECOVER_THREADS) { sR[i] = P.R[i]; sS[i] = P.S ? P.S[i] : 0.0; }
  __syncthreads();
  double* tX = tiles + warp * recover_small_warp_doubles<NX, NU, RCH>();
  double* tEX = tX + 32 * PX;
  double* tV = tEX + 32 * PX;
  double* tU = tV + 32 * PU;
  double* tEU = tU + 32 * PU;
  const long long pbase = ((long long)blockIdx.x * (RECOVER_THREADS / 32) + warp) * 32;
  if (pbase >= P.batch) return;
  const int nvalid = (int)((P.batch - pbase) < 32 ? (P.batch - pbase) : 32);
  const bool active = lane < nvalid;
  const long long p = pbase + (active ? lane : 0);
  // 16-byte accesses need even row lengths and 16-byte aligned bases (cudaMalloc'ed arrays are; offsets into them may not be)
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const bool vx = (NX % 2 == 0) && ((((long long)(H + 1) * NX) & 1) == 0) && al16(P.x) && al16(P.e_x);
  const bool vu = (NU % 2 == 0) && ((((long long)H * NU) & 1) == 0) && al16(P.u) && al16(P.e_u) && al16(P.v);
  double e[NX], xr[NX], ur[NU], up[NU];
#pragma unroll
  for (int i = 0; i < NX; i++) { xr[i] = P.xref[(P.xref_bc ? 0 : p) * NX + i]; e[i] = P.x0[p * NX + i] - xr[i]; }
#pragma unroll
  for (int i = 0; i < NU; i++) { ur[i] = P.uref[(P.uref_bc ? 0 : p) * NU + i]; up[i] = 0.0; }
  double J = 0.0;
  const int nchunks = (H + 1 + RCH - 1) / RCH;
  for (int c = 0; c < nchunks; c++) {
    const int k0 = c * RCH;
    const int ns_x = (H + 1 - k0) < RCH ? (H + 1 - k0) : RCH;        // states k0 .. k0 + ns_x - 1
    const int ns_u = (H - k0) < RCH ? (H - k0 > 0 ? H - k0 : 0) : RCH;
    if (ns_u > 0) warp_tile_copy<false>(tV, PU, const_cast<double*>(P.v) + pbase * (long long)NU * H + k0 * NU, (long long)NU * H, ns_u * NU, nvalid, lane, vu);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < RCH; q++) {
      const int k = k0 + q;
      if (k <= H) {
        const double* W = (k == H) ? sPt : sQ;
        double quad = 0.0;
#pragma unroll
        for (int i = 0; i < NX; i++) {
          tEX[lane * PX + q * NX + i] = e[i];
          tX[lane * PX + q * NX + i] = e[i] + xr[i];
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < NX; j++) s = fma(W[j * NX + i], e[j], s);
          quad = fma(e[i], s, quad);
        }
        J += quad;
      }
      if (k < H) {
        double uk[NU], eu[NU];
#pragma unroll
        for (int i = 0; i < NU; i++) {
          uk[i] = tV[lane * PU + q * NU + i]; eu[i] = uk[i] - ur[i];
          tU[lane * PU + q * NU + i] = uk[i]; tEU[lane * PU + q * NU + i] = eu[i];
          if (k == 0 && P.u0 && active) P.u0[p * NU + i] = uk[i];
        }
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
    __syncwarp();
    const long long gx = pbase * (long long)(H + 1) * NX + (long long)k0 * NX, gu = pbase * (long long)H * NU + (long long)k0 * NU;
    if (P.x) warp_tile_copy<true>(tX, PX, P.x + gx, (long long)(H + 1) * NX, ns_x * NX, nvalid, lane, vx);
    if (P.e_x) warp_tile_copy<true>(tEX, PX, P.e_x + gx, (long long)(H + 1) * NX, ns_x * NX, nvalid, lane, vx);
    if (ns_u > 0) {
      if (P.u) warp_tile_copy<true>(tU, PU, P.u + gu, (long long)H * NU, ns_u * NU, nvalid, lane, vu);
      if (P.e_u) warp_tile_copy<true>(tEU, PU, P.e_u + gu, (long long)H * NU, ns_u * NU, nvalid, lane, vu);
    }
    __syncwarp();
  }
  if (P.objective && active) P.objective[p] = J;
}

// Small batches (the closed-loop, few-plants-at-a-time use): the same per-problem arithmetic as recover_small_kernel -- expression by expression, so the
// results are bit-identical -- without the shared-memory tiles, chunking and warp-wide copies that make the tiled kernel coalesce for 10^4+ problems and cost
// it 21 us for ONE problem (ncu): one thread per problem, matrices through the read-only cache, results stored directly.
template <int NX, int NU>
__global__ void __launch_bounds__(64) recover_direct_kernel(const RecoverParams P) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.batch) return;
  const int H = P.H;
  double sA[NX * NX], sB[NX * NU], sQ[NX * NX], sPt[NX * NX], sR[NU * NU], sS[NU * NU];
#pragma unroll
  for (int i = 0; i < NX * NX; i++) { sA[i] = __ldg(P.A + i); sQ[i] = __ldg(P.Q + i); sPt[i] = __ldg(P.Pt + i); }
#pragma unroll
  for (int i = 0; i < NX * NU; i++) sB[i] = __ldg(P.B + i);
#pragma unroll
  for (int i = 0; i < NU * NU; i++) { sR[i] = __ldg(P.R + i); sS[i] = P.S ? __ldg(P.S + i) : 0.0; }
  double e[NX], xr[NX], ur[NU], up[NU];
#pragma unroll
  for (int i = 0; i < NX; i++) { xr[i] = P.xref[(P.xref_bc ? 0 : p) * NX + i]; e[i] = P.x0[p * NX + i] - xr[i]; }
#pragma unroll
  for (int i = 0; i < NU; i++) { ur[i] = P.uref[(P.uref_bc ? 0 : p) * NU + i]; up[i] = 0.0; }
  double J = 0.0;
  for (int k = 0; k <= H; k++) {
    {
      double quad = 0.0;
#pragma unroll
      for (int i = 0; i < NX; i++) {
        if (P.e_x) P.e_x[(p * (H + 1) + k) * NX + i] = e[i];
        if (P.x) P.x[(p * (H + 1) + k) * NX + i] = e[i] + xr[i];
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < NX; j++) s = fma((k == H) ? sPt[j * NX + i] : sQ[j * NX + i], e[j], s);
        quad = fma(e[i], s, quad);
      }
      J += quad;
    }
    if (k < H) {
      double uk[NU], eu[NU];
#pragma unroll
      for (int i = 0; i < NU; i++) {
        uk[i] = P.v[(p * H + k) * NU + i]; eu[i] = uk[i] - ur[i];
        if (P.u) P.u[(p * H + k) * NU + i] = uk[i];
        if (P.e_u) P.e_u[(p * H + k) * NU + i] = eu[i];
        if (k == 0 && P.u0) P.u0[p * NU + i] = uk[i];
      }
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
  if (P.objective) P.objective[p] = J;
}

inline bool recover_no_direct() { static const bool v = std::getenv("MPCB_NO_SMALL_COOP") != nullptr; return v; }      // A/B switch shared with the small-batch solve
constexpr long long RECOVER_DIRECT_MAX = 1184;      // batches up to here take the direct kernel (8 problems per SM: the small-batch regime of the solve)

// returns false when no specialisation exists (caller falls back to recover_kernel)
inline bool launch_recover_small(const RecoverParams& R, cudaStream_t st) {
  const unsigned grid = (unsigned)((R.batch + RECOVER_THREADS - 1) / RECOVER_THREADS);
  int dev_slot = 0;
  cudaGetDevice(&dev_slot);
  dev_slot &= 63;
#define MPCB_RS1(NX_, NU_, RCH_)                                                                                           \
  {                                                                                                                        \
    constexpr size_t smem = sizeof(double) * (RECOVER_THREADS / 32) * recover_small_warp_doubles<NX_, NU_, RCH_>();         \
    static bool attr_set[64] = {};      /* function attributes are per DEVICE (a multi-device handle launches on several) */  \
    if (smem > 48 * 1024 && !attr_set[dev_slot]) {                                                                        \
      cudaFuncSetAttribute(recover_small_kernel<NX_, NU_, RCH_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      attr_set[dev_slot] = true;                                                                                          \
    }                                                                                                                     \
    recover_small_kernel<NX_, NU_, RCH_><<<grid, RECOVER_THREADS, smem, st>>>(R);                                         \
    return true;                                                                                                          \
  }
#define MPCB_RS(NX_, NU_)                                                    \
  if (R.nx == NX_ && R.nu == NU_) {                                          \
    if (R.batch <= RECOVER_DIRECT_MAX && !recover_no_direct()) {             \
      recover_direct_kernel<NX_, NU_><<<(unsigned)((R.batch + 63) / 64), 64, 0, st>>>(R); \
      return true;                                                           \
    }                                                                        \
    if (grid > 444) MPCB_RS1(NX_, NU_, RECOVER_RCH_BATCH)                    \
    MPCB_RS1(NX_, NU_, RECOVER_RCH_FEW)                                      \
  }
  MPCB_RS(2, 1) MPCB_RS(2, 2) MPCB_RS(3, 1) MPCB_RS(3, 2) MPCB_RS(4, 1) MPCB_RS(4, 2) MPCB_RS(4, 4) MPCB_RS(6, 2) MPCB_RS(6, 3)
  MPCB_RS(8, 2) MPCB_RS(8, 4)
#undef MPCB_RS
#undef MPCB_RS1
  return false;
}

}  // namespace mpcb
