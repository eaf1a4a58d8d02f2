// Streamed batched ADMM for operators that do not fit the register file (nz + mg > 64); implementation and design
// notes in admm_stream.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/mpcb200.h"
#include "host_design.hpp"

namespace mpcb {

struct StreamConsts {
  double* T = nullptr;     // [NTp][NTp] row-major (symmetric), zero padded
  double* C = nullptr;     // [NTp][NTp]
  double* Lt = nullptr;    // [np][NTp]
  double* Lv = nullptr;    // [np][NTp] cold-start map (settings.cold_init) or null
  double* lo = nullptr;    // [NTp]
  double* hi = nullptr;
  double* rho = nullptr;
  double* rinv = nullptr;
  int NTp = 0;
};

struct StreamWork {
  // state (problem-major [rows][NTp]): Z holds c = (1-alpha) z + y/rho, Q holds q (box cols) / bound offset (general cols)
  double *X = nullptr, *Q = nullptr, *Z = nullptr, *YS = nullptr, *R0 = nullptr, *R1 = nullptr, *DY = nullptr;
  double *X2 = nullptr, *Q2 = nullptr, *Z2 = nullptr;             // compaction targets
  double *XT = nullptr, *YO = nullptr, *YP = nullptr;             // candidate x~ / y+ and y of the iteration before
  double *qn = nullptr, *qn2 = nullptr, *cert = nullptr;          // per-row |q|_inf, certificate partials [rows][3]
  unsigned long long *red = nullptr, *red2 = nullptr;             // [rows][4] residual reductions
  int *idx = nullptr, *idx2 = nullptr, *done = nullptr, *done2 = nullptr, *newly = nullptr;
  int* count = nullptr;        // device: [0] active rows after a check, [1] compaction cursor
  int* h_count = nullptr;      // pinned mirror
  int* sched = nullptr;        // device: ticket counter + per-(iteration, row block) completion counts of the fused-period kernel
  size_t sched_cap = 0;
  int resident_ctas = 0;       // CTAs of the iteration kernel that fit the device at once
  long long cap = 0;
  int NTp = 0;
};

struct StreamBatch {
  long long batch;
  const double *x0, *xref, *uref;
  int xref_bc, uref_bc;
  const double *warm_v, *warm_y;
  double *v_out, *y_out;
  int32_t *status, *iters;
  double *pres, *dres;
  // second rung of the rho ladder (mpcb_api.cu): row b of this solve is problem remap[b] of the caller's batch (null: b itself), and the
  // reported iteration counts continue from iters_add
  const int32_t* remap = nullptr;
  int iters_add = 0;
};

int stream_padded(int nt);
cudaError_t stream_upload(const Design& D, StreamConsts& sc, std::string& err);
cudaError_t stream_solve(const Design& D, const mpcb_settings& st, const StreamConsts& sc, StreamWork& sw, const StreamBatch& b,
                         int sm_count, cudaStream_t stream, int* launches, std::string& err);
void stream_release(StreamConsts& sc, StreamWork& sw);

}  // namespace mpcb
