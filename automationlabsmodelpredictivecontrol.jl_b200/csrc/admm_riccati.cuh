// Stage-wise ("sparse") batched ADMM for long horizons: the x-update  K x~ = r  of the condensed OSQP iteration is solved by a
// cached Riccati recursion over the horizon instead of a dense nz x nz contraction -- O(H) work and memory per iteration
// instead of O(H^2).  Implementation and design notes in admm_riccati.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mpcb200.h"

namespace mpcb {

constexpr int RIC_MAX_NX = 8, RIC_MAX_NU = 4;

struct RiccatiParams {
  // per-system constants
  const double* stage;   // [H][ric_stage_doubles(nx, nu)]: K, K', Lambda^-1, Acl, Acl' of every stage (host_design.cpp)
  const double* Lq;      // [nz][np] row-major: q(p) = Lq p
  const double* Lv;      // [nz][np] row-major cold-start map v_unc(p) = Lv p, or null (settings.cold_init = 0)
  double Bm[RIC_MAX_NX * RIC_MAX_NU];   // B  row-major [nx][nu]   (kernel parameters live in the constant bank: these are
  double Bt[RIC_MAX_NU * RIC_MAX_NX];   // B' row-major [nu][nx]    direct operands of the DFMAs, no load instruction)
  double lo[RIC_MAX_NU], hi[RIC_MAX_NU];
  int H, nz, np, ch;     // ch: stages per staged chunk
  double rho, sigma, alpha, eps_abs, eps_rel;
  int max_iter, check_every;
  // batch
  long long batch;
  const double* x0; const double* xref; const double* uref;
  int xref_bc, uref_bc;
  const double* warm_v; const double* warm_y;
  double* v_out; double* y_out;
  int32_t* status; int32_t* iters; double* pres; double* dres;
  // tile workspaces, [tiles][nz][32] doubles each (one warp owns one tile of 32 problems; row-major rows of 32 lanes)
  double* W; double* Qb; double* D; double* X; double* XT; double* YO;
  unsigned long long* counter;   // [0] tile queue head, [1] CTAs done (re-armed by the last CTA, like the on-chip kernel)
};

int ric_stage_doubles(int nx, int nu);
bool riccati_supported(int nx, int nu);
// picks warps per CTA and the chunk length for this (H, batch); returns false when the stage matrices do not fit shared memory
bool riccati_plan(int nx, int nu, int H, bool sig, long long batch, int sm_count, size_t smem_limit, int* warps_per_cta, int* ch, size_t* smem_bytes);
cudaError_t launch_riccati(int nx, int nu, bool sig, RiccatiParams P, int sm_count, size_t smem_limit, cudaStream_t st);

}  // namespace mpcb
