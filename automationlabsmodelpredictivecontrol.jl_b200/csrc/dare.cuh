// Device-side discrete algebraic Riccati equation for MANY systems (sm_100a): one warp per system.
//
// The reference computes the terminal weight of every controller it designs with
//   P_cost = ControlSystems.are(Discrete, A_sys, B_sys, Q, R)          (/root/reference/src/sub/design_mpc.jl:327)
// on the linearisation of the model at the reference (design_mpc.jl:319-323).  When every problem of a batch has its own
// reference -- hence its own linearisation -- that is one Riccati equation per problem.
// Here a warp runs the structured doubling algorithm on its system entirely in shared memory:
//   A_0 = A, G_0 = B R^-1 B', H_0 = Q;   W = I + G_k H_k;   A_{k+1} = A_k W^-1 A_k;
//   G_{k+1} = G_k + A_k W^-1 G_k A_k';   H_{k+1} = H_k + A_k' H_k W^-1 A_k;        H_k -> P  (quadratically)
// with W^-1 [A_k, G_k] from one Gauss-Jordan elimination with partial pivoting of the augmented matrix [W | A_k | G_k].
// Same recurrence and stopping rule as the host's dare_sda (host_design.cpp), which designs the single-system controllers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mpcb {

__host__ __device__ inline int dare_scratch_doubles(int nx) { return 8 * nx * nx; }

// C = X Y (TX: X' Y; TY: X Y'), all nx x nx column-major in shared memory; C must not alias X or Y.
template <bool TX, bool TY>
__device__ __forceinline__ void mm_warp(int nx, const double* X, const double* Y, double* Cm, int lane) {
  for (int o = lane; o < nx * nx; o += 32) {
    const int j = o / nx, i = o - j * nx;
    double s = 0.0;
    for (int k = 0; k < nx; k++) s = fma(TX ? X[i * nx + k] : X[k * nx + i], TY ? Y[k * nx + j] : Y[j * nx + k], s);
    Cm[o] = s;
  }
}

// A, B: the system (column-major, shared or global); Q (nx x nx), Rinv (nu x nu): weights; P: nx x nx output (shared);
// ws: dare_scratch_doubles(nx) of this warp's shared memory.  Returns the number of doubling steps, or -1 when the
// recurrence did not settle (singular pivot, non-finite values, or no convergence: (A, B) not stabilisable).
__device__ inline int dare_sda_warp(int nx, int nu, const double* A, const double* B, const double* Q, const double* Rinv, double* P, double* ws, int lane) {
  const int n2 = nx * nx;
  double* Ak = ws;            // nx x nx
  double* Gk = Ak + n2;
  double* Hk = Gk + n2;
  double* Wa = Hk + n2;       // nx x 3 nx augmented [W | A_k | G_k], column-major with nx rows
  double* T1 = Wa + 3 * n2;
  double* T2 = T1 + n2;
  // G_0 = B Rinv B'
  for (int o = lane; o < nx * nu; o += 32) {         // Wa = B Rinv  (nx x nu; fits the 3 nx^2 block for nu <= 3 nx, checked by the host)
    const int j = o / nx, i = o - j * nx;
    double s = 0.0;
    for (int k = 0; k < nu; k++) s = fma(B[k * nx + i], Rinv[j * nu + k], s);
    Wa[o] = s;
  }
  for (int o = lane; o < n2; o += 32) { Ak[o] = A[o]; Hk[o] = Q[o]; }
  __syncwarp();
  for (int o = lane; o < n2; o += 32) {
    const int j = o / nx, i = o - j * nx;
    double s = 0.0;
    for (int k = 0; k < nu; k++) s = fma(Wa[k * nx + i], B[k * nx + j], s);
    Gk[o] = s;
  }
  __syncwarp();
  for (int it = 1; it <= 100; it++) {
    // augmented matrix
    for (int o = lane; o < n2; o += 32) {
      const int j = o / nx, i = o - j * nx;
      double s = (i == j) ? 1.0 : 0.0;
      for (int k = 0; k < nx; k++) s = fma(Gk[k * nx + i], Hk[j * nx + k], s);
      Wa[o] = s; Wa[n2 + o] = Ak[o]; Wa[2 * n2 + o] = Gk[o];
    }
    __syncwarp();
    bool bad = false;
    for (int pv = 0; pv < nx; pv++) {
      // partial pivoting: row of the largest |W[i][pv]|, i >= pv
      double best = -1.0; int brow = pv;
      for (int i = pv + lane; i < nx; i += 32) { const double a = fabs(Wa[pv * nx + i]); if (a > best) { best = a; brow = i; } }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int orow = __shfl_xor_sync(0xffffffffu, brow, off);
        if (ob > best || (ob == best && orow < brow)) { best = ob; brow = orow; }
      }
      if (!(best > 1e-300) || !isfinite(best)) { bad = true; break; }
      if (brow != pv)
        for (int c = lane; c < 3 * nx; c += 32) { const double t = Wa[c * nx + pv]; Wa[c * nx + pv] = Wa[c * nx + brow]; Wa[c * nx + brow] = t; }
      __syncwarp();
      const double dinv = 1.0 / Wa[pv * nx + pv];
      __syncwarp();
      for (int c = lane; c < 3 * nx; c += 32) Wa[c * nx + pv] *= dinv;       // scaled pivot row (its pivot entry becomes 1)
      for (int i = lane; i < nx; i += 32) T1[i] = (i == pv) ? 0.0 : Wa[pv * nx + i];   // the pivot column, before it is eliminated
      __syncwarp();
      for (int o = lane; o < 3 * n2; o += 32) {
        const int c = o / nx, i = o - c * nx;
        Wa[o] = fma(-T1[i], Wa[c * nx + pv], Wa[o]);
      }
      __syncwarp();
    }
    if (bad) return -1;
    const double* V1 = Wa + n2;      // W^-1 A_k
    const double* V2 = Wa + 2 * n2;  // W^-1 G_k
    mm_warp<false, false>(nx, Ak, V2, T1, lane);     // T1 = A_k V2
    mm_warp<false, false>(nx, Hk, V1, T2, lane);     // T2 = H_k V1
    __syncwarp();
    double diff = 0.0, nrm = 0.0;
    for (int o = lane; o < n2; o += 32) {
      const int j = o / nx, i = o - j * nx;
      double sg = 0.0, sh = 0.0;
      for (int k = 0; k < nx; k++) { sg = fma(T1[k * nx + i], Ak[k * nx + j], sg); sh = fma(Ak[i * nx + k], T2[j * nx + k], sh); }   // T1 A_k' ; A_k' T2
      Gk[o] += sg;
      const double hn = Hk[o] + sh;
      diff = fmax(diff, fabs(sh)); nrm = fmax(nrm, fabs(hn));
      Hk[o] = hn;
    }
    __syncwarp();
    mm_warp<false, false>(nx, Ak, V1, T1, lane);     // A_{k+1} = A_k V1
    __syncwarp();
    for (int o = lane; o < n2; o += 32) Ak[o] = T1[o];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      diff = fmax(diff, __shfl_xor_sync(0xffffffffu, diff, off));
      nrm = fmax(nrm, __shfl_xor_sync(0xffffffffu, nrm, off));
    }
    __syncwarp();
    if (!isfinite(nrm) || !isfinite(diff)) return -1;
    if (diff <= 1e-15 * fmax(1.0, nrm)) {
      for (int o = lane; o < n2; o += 32) { const int j = o / nx, i = o - j * nx; P[o] = 0.5 * (Hk[o] + Hk[i * nx + j]); }
      __syncwarp();
      return it;
    }
  }
  return -1;
}

struct DareBatchParams {
  int nx, nu;
  long long batch;
  const double* A;      // [batch][nx*nx] column-major
  const double* B;      // [batch][nx*nu]
  const double* Q;      // nx x nx (shared by all systems)
  const double* Rinv;   // nu x nu
  double* P;            // [batch][nx*nx]
  int32_t* status;      // [batch]: doubling steps (> 0) or -1
};

constexpr int DARE_WARPS = 4;
__host__ __device__ inline size_t dare_batch_smem_bytes(int nx, int nu) {
  return sizeof(double) * ((size_t)nx * nx + (size_t)nu * nu + DARE_WARPS * ((size_t)dare_scratch_doubles(nx) + 2 * (size_t)nx * nx + (size_t)nx * nu));
}

static __global__ void __launch_bounds__(DARE_WARPS * 32) dare_batch_kernel(const DareBatchParams D) {
  extern __shared__ __align__(16) double dsm[];
  const int nx = D.nx, nu = D.nu, n2 = nx * nx;
  double* sQ = dsm;
  double* sRi = sQ + n2;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) sQ[i] = D.Q[i];
  for (int i = threadIdx.x; i < nu * nu; i += blockDim.x) sRi[i] = D.Rinv[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* w = sRi + nu * nu + (size_t)warp * (dare_scratch_doubles(nx) + 2 * n2 + nx * nu);
  double* sA = w; double* sB = sA + n2; double* sP = sB + nx * nu; double* ws = sP + n2;
  for (long long p = (long long)blockIdx.x * DARE_WARPS + warp; p < D.batch; p += (long long)gridDim.x * DARE_WARPS) {
    for (int o = lane; o < n2; o += 32) sA[o] = D.A[p * n2 + o];
    for (int o = lane; o < nx * nu; o += 32) sB[o] = D.B[p * nx * nu + o];
    __syncwarp();
    const int it = dare_sda_warp(nx, nu, sA, sB, sQ, sRi, sP, ws, lane);
    for (int o = lane; o < n2; o += 32) D.P[p * n2 + o] = it > 0 ? sP[o] : nan("");
    if (lane == 0 && D.status) D.status[p] = it;
    __syncwarp();
  }
}

}  // namespace mpcb
