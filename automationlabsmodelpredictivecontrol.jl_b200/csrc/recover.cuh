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

__host__ __device__ inline size_t recover_smem_bytes(int nx, int nu) {
  return sizeof(double) * ((size_t)(2 * nx + 2 * nu) * RECOVER_THREADS + 3 * nx * nx + nx * nu + 2 * nu * nu);
}

__global__ void __launch_bounds__(RECOVER_THREADS) recover_kernel(const RecoverParams P) {
  extern __shared__ __align__(16) double sm[];
  const int nx = P.nx, nu = P.nu, H = P.H, tid = threadIdx.x;
  double* se = sm;                                   // [nx][T] current deviation
  double* sn = se + nx * RECOVER_THREADS;            // [nx][T] next deviation
  double* su = sn + nx * RECOVER_THREADS;            // [nu][T] current input deviation
  double* sp = su + nu * RECOVER_THREADS;            // [nu][T] previous absolute input (S term)
  double* sA = sp + nu * RECOVER_THREADS;
  double* sB = sA + nx * nx;
  double* sQ = sB + nx * nu;
  double* sPt = sQ + nx * nx;
  double* sR = sPt + nx * nx;
  double* sS = sR + nu * nu;
  for (int i = tid; i < nx * nx; i += RECOVER_THREADS) { sA[i] = P.A[i]; sQ[i] = P.Q[i]; sPt[i] = P.Pt[i]; }
  for (int i = tid; i < nx * nu; i += RECOVER_THREADS) sB[i] = P.B[i];
  for (int i = tid; i < nu * nu; i += RECOVER_THREADS) { sR[i] = P.R[i]; sS[i] = P.S ? P.S[i] : 0.0; }
  __syncthreads();
  const long long p = (long long)blockIdx.x * RECOVER_THREADS + tid;
  if (p >= P.batch) return;
  const double* x0 = P.x0 + p * nx;
  const double* xr = P.xref + (P.xref_bc ? 0 : p) * nx;
  const double* ur = P.uref + (P.uref_bc ? 0 : p) * nu;
  const double* v = P.v + p * (long long)nu * H;
  for (int i = 0; i < nx; i++) se[i * RECOVER_THREADS + tid] = x0[i] - xr[i];
  double J = 0.0;
  for (int k = 0; k <= H; k++) {
    // write x_k, e_x_k and accumulate e' W e
    const double* W = (k == H) ? sPt : sQ;
    double quad = 0.0;
    for (int i = 0; i < nx; i++) {
      const double ei = se[i * RECOVER_THREADS + tid];
      if (P.e_x) P.e_x[(p * (H + 1) + k) * nx + i] = ei;
      if (P.x) P.x[(p * (H + 1) + k) * nx + i] = ei + xr[i];
      double s = 0.0;
      for (int j = 0; j < nx; j++) s = fma(W[j * nx + i], se[j * RECOVER_THREADS + tid], s);
      quad = fma(ei, s, quad);
    }
    J += quad;
    if (k == H) break;
    for (int i = 0; i < nu; i++) {
      const double ui = v[k * nu + i];
      const double eu = ui - ur[i];
      su[i * RECOVER_THREADS + tid] = eu;
      if (P.u) P.u[(p * H + k) * nu + i] = ui;
      if (P.e_u) P.e_u[(p * H + k) * nu + i] = eu;
      if (k == 0 && P.u0) P.u0[p * nu + i] = ui;
    }
    if (P.use_R) {
      double quadr = 0.0;
      for (int i = 0; i < nu; i++) {
        double s = 0.0;
        for (int j = 0; j < nu; j++) s = fma(sR[j * nu + i], su[j * RECOVER_THREADS + tid], s);
        quadr = fma(su[i * RECOVER_THREADS + tid], s, quadr);
      }
      J += quadr;
      if (P.use_S) {
        if (k > 0) {  // delta_u_{k-1} = u_{k-1} - u_k  (design_mpc.jl:429-432)
          double quads = 0.0;
          for (int i = 0; i < nu; i++) {
            double s = 0.0;
            for (int j = 0; j < nu; j++) s = fma(sS[j * nu + i], sp[j * RECOVER_THREADS + tid] - v[k * nu + j], s);
            quads = fma(sp[i * RECOVER_THREADS + tid] - v[k * nu + i], s, quads);
          }
          J += quads;
        }
        for (int i = 0; i < nu; i++) sp[i * RECOVER_THREADS + tid] = v[k * nu + i];
      }
    }
    for (int i = 0; i < nx; i++) {
      double s = 0.0;
      for (int j = 0; j < nx; j++) s = fma(sA[j * nx + i], se[j * RECOVER_THREADS + tid], s);
      for (int j = 0; j < nu; j++) s = fma(sB[j * nx + i], su[j * RECOVER_THREADS + tid], s);
      sn[i * RECOVER_THREADS + tid] = s;
    }
    double* tmp = se; se = sn; sn = tmp;
  }
  if (P.objective) P.objective[p] = J;
}

}  // namespace mpcb
