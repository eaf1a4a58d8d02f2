// Launchers of nmpc_sqp_kernel<ROWS, EQ, SB>: one translation unit per (EQ, SB) pair (nmpc_variant_*.cu) so the
// instantiations compile in parallel (nmpc_variant_lin_*.cu: the per-problem re-linearised LIN = true kernels).
#pragma once
#include <cuda_runtime.h>

#include "nmpc.cuh"

namespace mpcb {

cudaError_t launch_sqp_00(int rows, const NmpcParams& P, unsigned grid, int threads, size_t smem, size_t* smem_set, cudaStream_t st);
cudaError_t launch_sqp_01(int rows, const NmpcParams& P, unsigned grid, int threads, size_t smem, size_t* smem_set, cudaStream_t st);
cudaError_t launch_sqp_10(int rows, const NmpcParams& P, unsigned grid, int threads, size_t smem, size_t* smem_set, cudaStream_t st);
cudaError_t launch_sqp_11(int rows, const NmpcParams& P, unsigned grid, int threads, size_t smem, size_t* smem_set, cudaStream_t st);

cudaError_t launch_lin_00(int rows, const NmpcParams& P, unsigned grid, int threads, size_t smem, size_t* smem_set, cudaStream_t st);
cudaError_t launch_lin_01(int rows, const NmpcParams& P, unsigned grid, int threads, size_t smem, size_t* smem_set, cudaStream_t st);
cudaError_t launch_lin_10(int rows, const NmpcParams& P, unsigned grid, int threads, size_t smem, size_t* smem_set, cudaStream_t st);
cudaError_t launch_lin_11(int rows, const NmpcParams& P, unsigned grid, int threads, size_t smem, size_t* smem_set, cudaStream_t st);

// the per-problem re-linearised (LIN) variants
inline cudaError_t launch_lin(bool eq, bool sb, int rows, const NmpcParams& P, unsigned grid, int threads, size_t smem, size_t* smem_set, cudaStream_t st) {
  if (eq && sb) return launch_lin_11(rows, P, grid, threads, smem, smem_set, st);
  if (eq) return launch_lin_10(rows, P, grid, threads, smem, smem_set, st);
  if (sb) return launch_lin_01(rows, P, grid, threads, smem, smem_set, st);
  return launch_lin_00(rows, P, grid, threads, smem, smem_set, st);
}

inline cudaError_t launch_sqp(bool eq, bool sb, int rows, const NmpcParams& P, unsigned grid, int threads, size_t smem, size_t* smem_set, cudaStream_t st) {
  if (eq && sb) return launch_sqp_11(rows, P, grid, threads, smem, smem_set, st);
  if (eq) return launch_sqp_10(rows, P, grid, threads, smem, smem_set, st);
  if (sb) return launch_sqp_01(rows, P, grid, threads, smem, smem_set, st);
  return launch_sqp_00(rows, P, grid, threads, smem, smem_set, st);
}

}  // namespace mpcb
