// Closed-loop batched simulation kept on the GPU: the way the reference's controller is
// actually used is  update_initialization!(C, x) -> calculate!(C) -> apply u[:,1] -> measure the next x
// (/root/reference/src/main/computation_mpc.jl:17-55, pattern of test/computation_mpc_test.jl:94-103).  For a batch the
// whole loop stays on the device: solve kernel -> this plant-step kernel -> solve kernel (warm-started) -> ...
// The plant is the controller's own deviation model  x+ = x_ref + A (x - x_ref) + B (u0 - u_ref)  (linear.jl:59).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mpcb {

struct PlantStepParams {
  const double* A;      // nx x nx column-major
  const double* B;      // nx x nu
  int nx, nu, nz, t, steps;
  long long batch;
  const double* xref; const double* uref;
  int xref_bc, uref_bc;
  const double* v;      // [batch][nz] solution of this step (absolute inputs, stage-major)
  const int32_t* status; const int32_t* iters;
  const double* x_in;   // [batch][nx] x_t
  double* x_out;        // [batch][nx] x_{t+1}
  double* x_traj;       // [batch][steps+1][nx] or null
  double* u_traj;       // [batch][steps][nu] or null
  int32_t* iters_total; // [batch] or null
  int32_t* unsolved;    // [batch] or null: number of steps whose solve did not end with status 1
};

__global__ void plant_step_kernel(const PlantStepParams P) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.batch) return;
  const int nx = P.nx, nu = P.nu;
  const double* xr = P.xref + (P.xref_bc ? 0 : b) * nx;
  const double* ur = P.uref + (P.uref_bc ? 0 : b) * nu;
  const double* x = P.x_in + b * nx;
  const double* u0 = P.v + b * (long long)P.nz;
  if (P.t == 0 && P.x_traj)
    for (int i = 0; i < nx; i++) P.x_traj[(b * (P.steps + 1)) * nx + i] = x[i];
  if (P.u_traj)
    for (int i = 0; i < nu; i++) P.u_traj[(b * P.steps + P.t) * nu + i] = u0[i];
  for (int i = 0; i < nx; i++) {
    double s = xr[i];
    for (int j = 0; j < nx; j++) s = fma(P.A[j * nx + i], x[j] - xr[j], s);
    for (int j = 0; j < nu; j++) s = fma(P.B[j * nx + i], u0[j] - ur[j], s);
    P.x_out[b * nx + i] = s;
    if (P.x_traj) P.x_traj[(b * (P.steps + 1) + P.t + 1) * nx + i] = s;
  }
  if (P.iters_total) P.iters_total[b] = (P.t == 0 ? 0 : P.iters_total[b]) + P.iters[b];
  if (P.unsolved) P.unsolved[b] = (P.t == 0 ? 0 : P.unsolved[b]) + (P.status[b] == 1 ? 0 : 1);
}

}  // namespace mpcb
