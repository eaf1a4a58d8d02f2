// Nonlinear path (sm_100a): batched rollout of the Flux-style neural dynamics, their forward-mode Jacobians, and a
// Gauss-Newton SQP loop whose QP subproblem is solved by the same OSQP-style ADMM as the linear path.
//
// Replaces, for thousands of problems at once, what the reference does per problem with JuMP + Ipopt on the NL modelers
//   /root/reference/src/sub/model_modeler_implementation/fnn/mpc_modeler_implementation_fnn.jl:63-189
//   /root/reference/src/sub/model_modeler_implementation/resnet/mpc_modeler_implementation_resnet.jl:62-188
// (network layout: W_in without bias, n_hidden x [W_j, b_j] with activation -- resnet adds the skip --, W_out without
// bias; dynamics in absolute coordinates x_{k+1} = f(x_k, u_k); input box; cost of src/sub/design_mpc.jl:405-465) and the
// linearisation `proceed_system_linearization` used by the linear method on black-box models (fnn.jl:37-46).
//
// One WARP owns one problem.  All per-problem data (trajectory, sensitivities Gamma, the Gauss-Newton Hessian, its
// inverse, ADMM vectors) lives in that warp's slice of shared memory; the network weights and cost matrices are staged
// once per CTA.  Every matrix is stored so that the 32 lanes of a product read consecutive addresses (Julia's
// column-major weights are used as they come).  Per SQP iteration:
//   1. rollout + forward-mode Jacobians A_k, B_k;  Gamma_{k+1} = A_k Gamma_k (+ B_k in block k);
//      K += 2 Gamma' W Gamma,  g += 2 Gamma' W e  (W = Q, last stage P)              [Gauss-Newton condensed QP]
//   2. q = g - K u;  K += (sigma + rho) I;  K <- K^-1 in place (Gauss-Jordan, SPD, no pivoting)
//   3. ADMM in absolute inputs v (box umin <= v <= umax), warm-started at (u, y): per iteration one K^-1 mat-vec from
//      shared memory + the fused projection / dual update / residual reductions of admm_onchip.cuh (box-only form)
//   4. step d = v - u; accept if ||d||_inf <= tol, else Armijo backtracking on the TRUE cost (forward rollouts only)
// Problems are handed out through a global atomic counter (SQP iteration counts vary from 2 to the cap); the warps of a CTA
// advance in barrier-separated rounds of one SQP iteration so that they share the instruction cache (see the main loop).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dare.cuh"

namespace mpcb {

enum : int { NN_FNN = 0, NN_RESNET = 1, NN_POLYNET = 2, NN_DENSENET = 3 };
enum : int { ACT_RELU = 0, ACT_TANH = 1, ACT_SIGMOID = 2, ACT_SWISH = 3, ACT_IDENTITY = 4 };

struct NetDev {            // device pointers, Julia column-major
  const double* W_in;      // nn x nin
  const double* W_h;       // nh x (nn x nn)
  const double* b_h;       // nh x nn
  const double* W_out;     // nx x nn
  int arch, act, nx, nu, nn, nh, nin;
  // densenet: hidden layer l (1..nh) maps the concatenation of the l earlier blocks, W_l is nn x (l nn); W_out is nx x ((nh+1) nn)
  __host__ __device__ size_t wh_count() const { return arch == NN_DENSENET ? (size_t)nn * nn * nh * (nh + 1) / 2 : (size_t)nh * nn * nn; }
  __host__ __device__ size_t wout_count() const { return arch == NN_DENSENET ? (size_t)nx * nn * (nh + 1) : (size_t)nx * nn; }
  __host__ __device__ size_t weight_count() const { return (size_t)nn * nin + wh_count() + (size_t)nh * nn + wout_count(); }
};

// The weights of one network staged in shared memory.
struct NetSm {
  const double *W_in, *W_h, *b_h, *W_out;
  int arch, act, nx, nu, nn, nh, nin;
};

__device__ __forceinline__ NetSm stage_network(const NetDev& N, double* dst, int tid, int nthreads) {
  NetSm S;
  S.arch = N.arch; S.act = N.act; S.nx = N.nx; S.nu = N.nu; S.nn = N.nn; S.nh = N.nh; S.nin = N.nin;
  const int n1 = N.nn * N.nin, n2 = (int)N.wh_count(), n3 = N.nh * N.nn, n4 = (int)N.wout_count();
  for (int i = tid; i < n1; i += nthreads) dst[i] = N.W_in[i];
  for (int i = tid; i < n2; i += nthreads) dst[n1 + i] = N.W_h[i];
  for (int i = tid; i < n3; i += nthreads) dst[n1 + n2 + i] = N.b_h[i];
  for (int i = tid; i < n4; i += nthreads) dst[n1 + n2 + n3 + i] = N.W_out[i];
  S.W_in = dst; S.W_h = dst + n1; S.b_h = dst + n1 + n2; S.W_out = dst + n1 + n2 + n3;
  return S;
}

__device__ __forceinline__ void act_eval(int id, double h, double& a, double& da) {
  switch (id) {
    case ACT_RELU: a = h > 0.0 ? h : 0.0; da = h > 0.0 ? 1.0 : 0.0; break;     // NNlib.relu, derivative 0 at 0
    case ACT_TANH: { const double t = tanh(h); a = t; da = 1.0 - t * t; break; }
    case ACT_SIGMOID: { const double s = 1.0 / (1.0 + exp(-h)); a = s; da = s * (1.0 - s); break; }
    case ACT_SWISH: { const double s = 1.0 / (1.0 + exp(-h)); a = h * s; da = s + h * s * (1.0 - s); break; }
    default: a = h; da = 1.0; break;
  }
}

// One layer of the forward-mode Jacobian:  Jw[c][i] = sdv[i] * sum_j W[i][j] Jr[c][j]  (+ addA[c][i]) (+ addB[c][i]),
// over flattened (input c, neuron i) pairs so that all 32 lanes work even for 13-neuron layers, two pairs per lane and pass.
__device__ __forceinline__ void jac_layer(const double* __restrict__ W, const double* __restrict__ Jr, double* __restrict__ Jw,
                                          const double* __restrict__ sdv, const double* addA, const double* addB, int nn, int nin, int lane) {
  const int tot = nn * nin;
  for (int o0 = lane; o0 < tot; o0 += 64) {
    const int o1 = o0 + 32;
    const bool has1 = o1 < tot;
    const int c0 = o0 / nn, i0 = o0 - c0 * nn;
    const int c1 = has1 ? o1 / nn : c0, i1 = has1 ? o1 - c1 * nn : i0;
    double t0 = 0.0, t1 = 0.0;
#pragma unroll 4
    for (int j = 0; j < nn; j++) {
      t0 = fma(W[j * nn + i0], Jr[c0 * nn + j], t0);
      t1 = fma(W[j * nn + i1], Jr[c1 * nn + j], t1);
    }
    t0 *= sdv[i0]; t1 *= sdv[i1];
    if (addA) { t0 += addA[o0]; if (has1) t1 += addA[o1]; }
    if (addB) { t0 += addB[o0]; if (has1) t1 += addB[o1]; }
    Jw[o0] = t0;
    if (has1) Jw[o1] = t1;
  }
}

// DenseNet (densenet/mpc_modeler_implementation_densenet.jl:128-162): y_1 = W_in [x;u];  y_j = [act(W_j y_{j-1} + b_{j-1}); y_{j-1}]
// (the new block goes in FRONT, so y_j has j nn entries);  x+ = W_out y_{nh+1}.  Blocks are kept in creation order in
// ycat[(nh+1) nn] (block 0 = y_1); column block bp of W_j multiplies creation block (l - 1 - bp).  JAC: Jcat[(nh+1)][nin][nn].
// __noinline__: kept out of the SQP kernel's register allocation (inlined it cost the common architectures spills in their hot loops)
template <bool JAC>
__device__ __noinline__ void nn_eval_dense_warp(const NetSm& N, const double* xu, double* f, double* ycat, double* Jcat, double* AB, double* sd,
                                                   int lane) {
  const int nn = N.nn, nin = N.nin, nx = N.nx, nh = N.nh;
  for (int i = lane; i < nn; i += 32) {
    double s = 0.0;
    for (int j = 0; j < nin; j++) s = fma(N.W_in[j * nn + i], xu[j], s);
    ycat[i] = s;
  }
  if (JAC)
    for (int i = lane; i < nn * nin; i += 32) Jcat[i] = N.W_in[i];
  __syncwarp();
  const double* W = N.W_h;
  for (int l = 1; l <= nh; l++) {
    const double* b = N.b_h + (l - 1) * nn;
    for (int i = lane; i < nn; i += 32) {
      double s = b[i];
      for (int bp = 0; bp < l; bp++) {
        const double* yb_ = ycat + (l - 1 - bp) * nn;
        const double* Wb = W + (size_t)bp * nn * nn;
#pragma unroll 4
        for (int j = 0; j < nn; j++) s = fma(Wb[j * nn + i], yb_[j], s);
      }
      double a, da;
      act_eval(N.act, s, a, da);
      ycat[l * nn + i] = a;
      if (JAC) sd[i] = da;
    }
    __syncwarp();
    if (JAC) {
      double* Jw = Jcat + (size_t)l * nn * nin;
      for (int o = lane; o < nn * nin; o += 32) {
        const int c = o / nn, i = o - c * nn;
        double t = 0.0;
        for (int bp = 0; bp < l; bp++) {
          const double* Jr = Jcat + (size_t)(l - 1 - bp) * nn * nin + c * nn;
          const double* Wb = W + (size_t)bp * nn * nn;
#pragma unroll 4
          for (int j = 0; j < nn; j++) t = fma(Wb[j * nn + i], Jr[j], t);
        }
        Jw[o] = t * sd[i];
      }
      __syncwarp();
    }
    W += (size_t)l * nn * nn;
  }
  for (int i = lane; i < nx; i += 32) {
    double s = 0.0;
    for (int bp = 0; bp <= nh; bp++) {
      const double* yb_ = ycat + (nh - bp) * nn;
      const double* Wb = N.W_out + (size_t)bp * nn * nx;
#pragma unroll 4
      for (int j = 0; j < nn; j++) s = fma(Wb[j * nx + i], yb_[j], s);
    }
    f[i] = s;
  }
  if (JAC)
    for (int o = lane; o < nx * nin; o += 32) {
      const int c = o / nx, i = o - c * nx;
      double s = 0.0;
      for (int bp = 0; bp <= nh; bp++) {
        const double* Jr = Jcat + (size_t)(nh - bp) * nn * nin + c * nn;
        const double* Wb = N.W_out + (size_t)bp * nn * nx;
#pragma unroll 4
        for (int j = 0; j < nn; j++) s = fma(Wb[j * nx + i], Jr[j], s);
      }
      AB[o] = s;
    }
  __syncwarp();
}

// Warp-cooperative network evaluation.  xu[nin] (shared) -> f[nx] (shared).  Scratch: ya, yb, yc3 [nn] (yc3: the PolyNet
// branch).  With JAC also AB[nin][nx] (column-major nx x nin: d f / d [x;u]) using Ja, Jb, Jc3 [nin][nn] and sd, sd2 [nn]
// (activation derivatives).  Ends with a __syncwarp().
//   fnn     y+ = act(W y + b)                                     (fnn.jl:133-141)
//   resnet  y+ = y + act(W y + b)                                 (resnet.jl:131-140)
//   polynet br = act(W y + b);  y+ = y + br + act(W br + b)       (polynet.jl:132-149: both paths share W_j and b_j)
template <bool JAC>
__device__ __forceinline__ void nn_eval_warp(const NetSm& N, const double* xu, double* f, double* ya, double* yb, double* Ja, double* Jb,
                                             double* AB, double* sd, int lane) {
  const int nn = N.nn, nin = N.nin, nx = N.nx;
  if (N.arch == NN_DENSENET) {     // the caller's scratch is one block [xu | f | ya ...]: the concatenated layers live from ya on
    double* ycat = ya;
    double* Jcat = ycat + (size_t)(N.nh + 1) * nn;
    double* ABd = Jcat + (JAC ? (size_t)(N.nh + 1) * nn * nin : 0);
    double* sdd = ABd + (JAC ? nx * nin : 0);
    nn_eval_dense_warp<JAC>(N, xu, f, ycat, Jcat, ABd, sdd, lane);
    if (JAC) {                     // hand the Jacobian back where the caller expects it
      for (int o0 = 0; o0 < nx * nin; o0 += 32) {          // the two regions may overlap: read a pass, then write it
        const int o = o0 + lane;
        const double v = o < nx * nin ? ABd[o] : 0.0;
        __syncwarp();
        if (o < nx * nin) AB[o] = v;
        __syncwarp();
      }
    }
    return;
  }
  double* yc3 = sd + (JAC ? 2 * nn : 0);            // scratch layout: [sd | sd2 |] br [| Jbr]   (see nn_eval_scratch_doubles)
  double* sd2 = sd + nn;
  double* Jc3 = yc3 + nn;
  for (int i = lane; i < nn; i += 32) {
    double s = 0.0;
#pragma unroll 6
    for (int j = 0; j < nin; j++) s = fma(N.W_in[j * nn + i], xu[j], s);
    ya[i] = s;
  }
  if (JAC)
    for (int i = lane; i < nn * nin; i += 32) Ja[i] = N.W_in[i];
  __syncwarp();
  double *yc = ya, *yn = yb, *Jc = Ja, *Jn = Jb;
  for (int l = 0; l < N.nh; l++) {
    const double* __restrict__ W = N.W_h + (size_t)l * nn * nn;
    const double* __restrict__ b = N.b_h + l * nn;
    for (int i = lane; i < nn; i += 32) {
      double s = b[i];
#pragma unroll 8
      for (int j = 0; j < nn; j++) s = fma(W[j * nn + i], yc[j], s);
      double a, da;
      act_eval(N.act, s, a, da);
      if (N.arch == NN_POLYNET) yc3[i] = a; else yn[i] = N.arch == NN_RESNET ? a + yc[i] : a;
      if (JAC) sd[i] = da;
    }
    __syncwarp();
    if (N.arch == NN_POLYNET) {
      for (int i = lane; i < nn; i += 32) {
        double s = b[i];
#pragma unroll 8
        for (int j = 0; j < nn; j++) s = fma(W[j * nn + i], yc3[j], s);
        double a, da;
        act_eval(N.act, s, a, da);
        yn[i] = yc[i] + yc3[i] + a;
        if (JAC) sd2[i] = da;
      }
      if (JAC) {
        jac_layer(W, Jc, Jc3, sd, nullptr, nullptr, nn, nin, lane);      // J_br = D1 W J_y
        __syncwarp();
        jac_layer(W, Jc3, Jn, sd2, Jc, Jc3, nn, nin, lane);              // J_y+ = J_y + J_br + D2 W J_br
      }
      __syncwarp();
    } else if (JAC) {
      jac_layer(W, Jc, Jn, sd, N.arch == NN_RESNET ? Jc : nullptr, nullptr, nn, nin, lane);
      __syncwarp();
    }
    double* tp = yc; yc = yn; yn = tp;
    if (JAC) { tp = Jc; Jc = Jn; Jn = tp; }
  }
  for (int i = lane; i < nx; i += 32) {
    double s = 0.0;
#pragma unroll 8
    for (int j = 0; j < nn; j++) s = fma(N.W_out[j * nx + i], yc[j], s);
    f[i] = s;
  }
  if (JAC)
    for (int o = lane; o < nx * nin; o += 32) {
      const int c = o / nx, i = o - c * nx;
      double s = 0.0;
#pragma unroll 8
      for (int j = 0; j < nn; j++) s = fma(N.W_out[j * nx + i], Jc[c * nn + j], s);
      AB[o] = s;
    }
  __syncwarp();
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) { const double w = __shfl_xor_sync(0xffffffffu, v, o); v = v > w ? v : w; }
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Phase profiling for development (-DMPCB_NMPC_PROF): per-phase clock64() totals of lane 0, accumulated in a global array.
#ifdef MPCB_NMPC_PROF
__device__ unsigned long long g_nmpc_prof[8];
#define NMPC_PROF_DECL long long _pt = clock64();
#define NMPC_PROF(slot) do { const long long _n = clock64(); if (lane == 0) atomicAdd(&g_nmpc_prof[slot], (unsigned long long)(_n - _pt)); _pt = _n; } while (0)
#else
#define NMPC_PROF_DECL
#define NMPC_PROF(slot)
#endif

constexpr int NN_WARPS = 4;
constexpr int NN_THREADS = NN_WARPS * 32;

// ---------------------------------------------------------------------------------------------------------------------
// mpcb_nn_rollout_batch / mpcb_nn_jacobian_batch
// ---------------------------------------------------------------------------------------------------------------------
struct NnBatchParams {
  NetDev net;
  long long batch;
  int H;
  const double* x0;   // [batch][nx]       (jacobian: the evaluation points x)
  const double* u;    // [batch][H][nu]    (jacobian: [batch][nu])
  double* x;          // [batch][H+1][nx]  (jacobian: f [batch][nx], may be null)
  double* A;          // [batch][nx x nx column-major]
  double* B;          // [batch][nx x nu column-major]
};

__host__ __device__ inline size_t nn_eval_scratch_doubles(const NetDev& N, bool jac) {
  const size_t std_ = (size_t)N.nin + N.nx + 3 * N.nn + (jac ? (size_t)3 * N.nn * N.nin + (size_t)N.nx * N.nin + 2 * N.nn : 0);
  const size_t dense = (size_t)N.nin + N.nx + (size_t)(N.nh + 1) * N.nn + (jac ? (size_t)(N.nh + 1) * N.nn * N.nin + 2 * (size_t)N.nx * N.nin + N.nn : 0);
  return N.arch == NN_DENSENET ? (dense > std_ ? dense : std_) : std_;
}
__host__ __device__ inline size_t nn_batch_smem_bytes(const NetDev& N, bool jac) {
  return sizeof(double) * (N.weight_count() + NN_WARPS * nn_eval_scratch_doubles(N, jac));
}

template <bool JAC>
__global__ void __launch_bounds__(NN_THREADS) nn_batch_kernel(const NnBatchParams P) {
  extern __shared__ __align__(16) double sm[];
  const NetSm N = stage_network(P.net, sm, threadIdx.x, NN_THREADS);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nx = N.nx, nu = N.nu, nin = N.nin, nn = N.nn;
  double* w = sm + P.net.weight_count() + warp * nn_eval_scratch_doubles(P.net, JAC);
  double* xu = w; double* f = xu + nin; double* ya = f + nx; double* yb = ya + nn;
  double* Ja = yb + nn; double* Jb = Ja + (JAC ? nn * nin : 0); double* AB = Jb + (JAC ? nn * nin : 0); double* sd = AB + (JAC ? nx * nin : 0);
  for (long long p = (long long)blockIdx.x * NN_WARPS + warp; p < P.batch; p += (long long)gridDim.x * NN_WARPS) {
    if (JAC) {
      for (int i = lane; i < nx; i += 32) xu[i] = P.x0[p * nx + i];
      for (int i = lane; i < nu; i += 32) xu[nx + i] = P.u[p * nu + i];
      __syncwarp();
      nn_eval_warp<true>(N, xu, f, ya, yb, Ja, Jb, AB, sd, lane);
      if (P.x) for (int i = lane; i < nx; i += 32) P.x[p * nx + i] = f[i];
      for (int o = lane; o < nx * nx; o += 32) P.A[p * nx * nx + o] = AB[o];
      for (int o = lane; o < nx * nu; o += 32) P.B[p * nx * nu + o] = AB[nx * nx + o];
      __syncwarp();
    } else {
      for (int i = lane; i < nx; i += 32) { const double v = P.x0[p * nx + i]; xu[i] = v; P.x[p * (P.H + 1) * nx + i] = v; }
      for (int k = 0; k < P.H; k++) {
        for (int i = lane; i < nu; i += 32) xu[nx + i] = P.u[(p * P.H + k) * nu + i];
        __syncwarp();
        nn_eval_warp<false>(N, xu, f, ya, yb, nullptr, nullptr, nullptr, sd, lane);
        for (int i = lane; i < nx; i += 32) { const double v = f[i]; xu[i] = v; P.x[(p * (P.H + 1) + k + 1) * nx + i] = v; }
        __syncwarp();
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// mpcb_solve_nmpc_batch: the SQP kernel
// ---------------------------------------------------------------------------------------------------------------------
struct NmpcParams {
  NetDev net;
  const double* Q;    // nx x nx
  const double* Pt;   // nx x nx terminal weight
  const double* Hc;   // nz x nz constant Hessian part 2(I (x) R) + 2 D'(I (x) S) D   (design_mpc.jl:436-447)
  const double* lb;   // nz
  const double* ub;   // nz
  int H, nz;
  double rho, sigma, alpha, eps_abs, eps_rel;
  int max_iter, check_every;
  double sqp_tol, ls_c1, ls_noise;
  double rho_eq_scale;   // terminal equality rows: rho_e,i = rho_eq_scale * rho / |Gamma_H,i|^2
  int sync_rounds;       // 1: the warps of a CTA advance in barrier-separated rounds of one SQP iteration (see the kernel's main loop)
  int term_ball;         // EQ kernels: the terminal rows form the contractive ball |e_H|_2 <= sqrt(0.9) |e_0|_2 (design_mpc.jl:333-340) instead of e_H = 0
  const double* Rinv;    // nu x nu (LIN kernels: G_0 = B R^-1 B' of the per-problem Riccati equation)
  int lin_dare;          // LIN kernels: 1 = terminal weight from the per-problem DARE, 0 = the design's Pt for every problem
  const double* xmin;    // nx, state box (SB kernels only)
  const double* xmax;
  int sqp_max_iter, ls_max;
  long long batch;
  const double* x0; const double* xref; const double* uref;
  int xref_bc, uref_bc;
  const double* warm_u;   // [batch][nz] initial guess (null: the reference input clipped to the box)
  const double* warm_y;   // [batch][nz (+ nx with the terminal equality)] duals of the box rows [and terminal rows] (null: 0)
  double *u, *e_u, *x, *e_x, *u0, *objective, *y;   // outputs, any may be null
  int32_t *status, *iters, *inner_iters;
  double *step, *qp_dres;
  unsigned long long* counter;
};

__host__ __device__ inline int nmpc_ldk(int nz) { return nz | 1; }
// per-warp shared memory (doubles)
__host__ __device__ inline int nmpc_ldg(int nz, bool sb) { return sb ? (nz | 1) : nz; }   // odd pitch: row-per-lane dot products are conflict-free
__host__ __device__ inline size_t nmpc_warp_doubles(const NetDev& N, int H, int nz, bool sb, bool lin = false) {
  return (lin ? (size_t)N.nx * N.nin + (size_t)N.nx * N.nx + dare_scratch_doubles(N.nx) : 0)      // LIN: [A B] and P of this problem, Riccati workspace
         + (size_t)nz * nmpc_ldk(nz)            // K
         + (sb ? (size_t)(H + 1) * N.nx * nmpc_ldg(nz, true)      // all Gamma_k (state-box rows) + W Gamma scratch
               : 2 * (size_t)N.nx * nz)       // Gamma double buffer (the idle one holds W Gamma)
         + (sb ? 8 * (size_t)H * N.nx : 0)    // state-box rows: rho_g, lo, hi, z_g, ys_g, m, n, y_g
         + 6 * (size_t)nz                     // u, v, r, g, q, col
         + (size_t)(H + 1) * N.nx             // trajectory
         + 2 * (size_t)N.nx                   // e, We
         + 8 * (size_t)N.nx                   // terminal equality rows: rho_e, b, z_g, ys_g, m, n, t_g, y_g
         + nn_eval_scratch_doubles(N, true);
}
__host__ __device__ inline size_t nmpc_const_doubles(const NetDev& N, int nz) {
  return N.weight_count() + 2 * (size_t)N.nx * N.nx + (size_t)nz * nz + 2 * (size_t)nz + 2 * (size_t)N.nx;
}
__host__ __device__ inline size_t nmpc_smem_bytes(const NetDev& N, int H, int nz, int warps, bool sb, bool lin = false) {
  return sizeof(double) * (nmpc_const_doubles(N, nz) + (size_t)warps * nmpc_warp_doubles(N, H, nz, sb, lin));
}

// ROWS = ceil(nz / 32): decision-variable rows owned by each lane (row e = lane + 32 * i)
constexpr int NMPC_MAX_WARPS = 10;      // CTA width is chosen at design time to maximise resident warps per SM (shared-memory bound)
// EQ: terminal equality e_x[:,end] == 0 (design_mpc.jl:330-331) as nx linearised rows Gamma_H v = Gamma_H u - e_H(u); with
//     P.term_ball the same rows carry the contractive set e_H' e_H <= 0.9 e_0' e_0 (design_mpc.jl:333-340) linearised as
//     |Gamma_H v - b|_2 <= rad: the ADMM projects them onto that ball (one common step size) instead of the point b.
// SB: state box xmin <= x[:,k] <= xmax for k = 2..H+1 (fnn.jl:146-154) as nx*H linearised inequality rows.
// LIN: the reference's LINEAR method on a black-box model (design_mpc.jl:319-327), re-designed per problem on the device:
//      the network is linearised once at this problem's reference (x_ref, u_ref), the terminal weight comes from this
//      problem's Riccati equation (dare.cuh), the prediction model is the deviation model x+ = x_ref + A (x - x_ref) +
//      B (u - u_ref) of linear.jl:59, and the (now exact) condensed QP is solved once.  Everything else is shared.
template <int ROWS, bool EQ, bool SB, bool LIN>
__global__ void __launch_bounds__(NMPC_MAX_WARPS * 32, 1) nmpc_sqp_kernel(const NmpcParams P) {
  extern __shared__ __align__(16) double sm[];
  const int nwarps = blockDim.x >> 5;
  const NetSm N = stage_network(P.net, sm, threadIdx.x, blockDim.x);
  const int nx = N.nx, nu = N.nu, nin = N.nin, nn = N.nn, H = P.H, nz = P.nz, ldk = nmpc_ldk(nz), ldg = nmpc_ldg(nz, SB), ms = SB ? nx * H : 0;
  double* sQ = sm + P.net.weight_count();
  double* sPt = sQ + nx * nx;
  double* sHc = sPt + nx * nx;
  double* sLb = sHc + (size_t)nz * nz;
  double* sUb = sLb + nz;
  double* sXmin = sUb + nz;
  double* sXmax = sXmin + nx;
  for (int i = threadIdx.x; i < nx * nx; i += blockDim.x) { sQ[i] = P.Q[i]; sPt[i] = P.Pt[i]; }
  for (int i = threadIdx.x; i < nz * nz; i += blockDim.x) sHc[i] = P.Hc[i];
  for (int i = threadIdx.x; i < nz; i += blockDim.x) { sLb[i] = P.lb[i]; sUb[i] = P.ub[i]; }
  for (int i = threadIdx.x; i < nx; i += blockDim.x) { sXmin[i] = SB ? P.xmin[i] : 0.0; sXmax[i] = SB ? P.xmax[i] : 0.0; }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  (void)nwarps;
  double* w = sXmax + nx + (size_t)warp * nmpc_warp_doubles(P.net, H, nz, SB, LIN);
  double* lAB = w;                       w += LIN ? nx * nin : 0;      // LIN: [A B] at this problem's reference
  double* lP = w;                        w += LIN ? nx * nx : 0;       //      terminal weight of this problem
  double* lws = w;                       w += LIN ? dare_scratch_doubles(nx) : 0;
  double* K = w;                         w += (size_t)nz * ldk;
  double* G0 = w;                        w += SB ? (size_t)H * nx * ldg : (size_t)nx * nz;   // SB: Gamma_1 .. Gamma_H, one block of nx rows each
  double* G1 = w;                        w += SB ? (size_t)nx * ldg : (size_t)nx * nz;       // SB: W Gamma scratch
  double* grho = w;                      w += ms;     // SB only, per state-box row: step size
  double* glo = w;                       w += ms;     //   bounds of the linearised row
  double* ghi = w;                       w += ms;
  double* gz = w;                        w += ms;     //   z
  double* gys = w;                       w += ms;     //   y / rho
  double* gm = w;                        w += ms;     //   rho (z - ys)
  double* gn = w;                        w += ms;     //   rho (ys+ - t_g)
  double* gy = w;                        w += ms;     //   multipliers carried across SQP iterations
  double* su = w;                        w += nz;     // current iterate u
  double* sv = w;                        w += nz;     // QP solution / line-search candidate
  double* sr = w;                        w += nz;     // ADMM right-hand side
  double* sg = w;                        w += nz;     // gradient
  double* sq = w;                        w += nz;     // q = g - K u
  double* scol = w;                      w += nz;     // Gauss-Jordan pivot column
  double* traj = w;                      w += (size_t)(H + 1) * nx;
  double* se = w;                        w += nx;
  double* sWe = w;                       w += nx;
  double* srhoe = w;                     w += nx;     // EQ only: per-row step size
  double* sb = w;                        w += nx;     // right-hand side of the linearised terminal rows
  double* szg = w;                       w += nx;     // z of the terminal rows
  double* sysg = w;                      w += nx;     // y / rho_e of the terminal rows
  double* ssm = w;                       w += nx;     // rho_e (z_g - ys_g): operand of the G' product in the right-hand side
  double* ssn = w;                       w += nx;     // rho_e (ys_g+ - t_g): operand of the G' product in the dual residual
  double* stg = w;                       w += nx;     // G x~
  double* syg = w;                       w += nx;     // multipliers of the terminal rows (carried across SQP iterations)
  double* xu = w; double* f = xu + nin; double* ya = f + nx; double* yb = ya + nn;
  double* Ja = yb + nn; double* Jb = Ja + nn * nin; double* AB = Jb + nn * nin; double* sd = AB + nx * nin;

  const double rho = P.rho, sigma = P.sigma, alpha = P.alpha, oma = 1.0 - P.alpha, sig_rho = P.sigma + P.rho;
  const int max_iter = ((P.max_iter + P.check_every - 1) / P.check_every) * P.check_every;

  // Per-problem state of this warp, kept across rounds (see the loop below).
  long long p = -1;         // problem held by this warp (-1: none)
  bool exhausted = false;   // the work queue has run dry
  const double *x0 = nullptr, *xr = nullptr, *ur = nullptr;
  double rad = 0.0;
  const double* Wt = sPt;   // terminal weight
  const double* ABk = AB;   // stage Jacobians [A_k B_k]
  double yd[ROWS];          // duals of the box rows (carried across SQP iterations)
  double mu = 0.0;          // l1 merit weight of the terminal rows
  const int ny = nz + ms + (EQ ? nx : 0);       // duals: [input box | state-box rows | terminal rows]
  int status = -2, sqp_it = 0, inner_total = 0;
  double step = 0.0, qp_rd = 0.0, Jcur = 0.0;
  bool have_traj = false;
#pragma unroll
  for (int i = 0; i < ROWS; i++) yd[i] = 0.0;
  // one step of the prediction model from xu = [x; u]: the network, or (LIN) the deviation model around the reference
  auto model_step = [&]() {
    if (LIN) {
      for (int i = lane; i < nx; i += 32) {
        double sacc = xr[i];
        for (int j = 0; j < nx; j++) sacc = fma(lAB[j * nx + i], xu[j] - xr[j], sacc);
        for (int j = 0; j < nu; j++) sacc = fma(lAB[(nx + j) * nx + i], xu[nx + j] - ur[j], sacc);
        f[i] = sacc;
      }
      __syncwarp();
    } else {
      nn_eval_warp<false>(N, xu, f, ya, yb, nullptr, nullptr, nullptr, sd, lane);
    }
  };
  // forward rollout of the inputs in `uu` -> traj, returns the true cost J (all lanes)
  auto rollout_cost = [&](const double* uu) -> double {
    double J = 0.0;
    for (int i = lane; i < nx; i += 32) { const double v = x0[i]; xu[i] = v; traj[i] = v; }
    __syncwarp();
    for (int k = 0; k <= H; k++) {
      const double* W = (k == H) ? Wt : sQ;
      double part = 0.0;
      for (int i = lane; i < nx; i += 32) {
        double s = 0.0;
        for (int j = 0; j < nx; j++) s = fma(W[j * nx + i], xu[j] - xr[j], s);
        part = fma(xu[i] - xr[i], s, part);
      }
      J += part;
      if (k == H) break;
      for (int i = lane; i < nu; i += 32) xu[nx + i] = uu[k * nu + i];
      __syncwarp();
      model_step();
      for (int i = lane; i < nx; i += 32) { const double v = f[i]; xu[i] = v; traj[(k + 1) * nx + i] = v; }
      __syncwarp();
    }
    double part = 0.0;                 // 1/2 du' Hc du
    for (int e = lane; e < nz; e += 32) scol[e] = uu[e] - ur[e % nu];
    __syncwarp();
    for (int e = lane; e < nz; e += 32) {
      double s = 0.0;
#pragma unroll 8
      for (int j = 0; j < nz; j++) s = fma(sHc[j * nz + e], scol[j], s);
      part = fma(0.5 * scol[e], s, part);
    }
    __syncwarp();
    return warp_sum(J + part);
  };

  // The warps of a CTA advance in ROUNDS of one SQP iteration each, separated by a CTA barrier: a warp whose problem finished
  // fetches the next one at the round boundary.  In step, the warps execute the same few KB of this large kernel at the same
  // time and share the instruction cache; left alone they drift apart within a few problems and the cache thrashes (measured
  // on the re-linearised solve: 65 536 problems 28.7 -> 19.4 ms; SQP on the ResNet surrogate 45.5 -> 34.3 ms; a second barrier after
  // the variable-length ADMM phase changes nothing).
  NMPC_PROF_DECL
  while (true) {
    if (p < 0 && !exhausted) {
      long long pn = 0;
      if (lane == 0) pn = (long long)atomicAdd(P.counter, 1ULL);
      pn = __shfl_sync(0xffffffffu, pn, 0);
      if (pn >= P.batch) exhausted = true;
      else {
        p = pn;
        x0 = P.x0 + p * nx;
        xr = P.xref + (P.xref_bc ? 0 : p) * nx;
        ur = P.uref + (P.uref_bc ? 0 : p) * nu;
        rad = 0.0;                // contractive terminal set: radius sqrt(0.9) |x0 - xref|_2 of this problem's ball
        if (EQ && P.term_ball) {
          double d2 = 0.0;
          for (int i = 0; i < nx; i++) { const double dv = x0[i] - xr[i]; d2 = fma(dv, dv, d2); }
          rad = sqrt(0.9 * d2);
        }
        Wt = sPt; ABk = AB;
        if (LIN) {
          for (int i = lane; i < nx; i += 32) xu[i] = xr[i];
          for (int i = lane; i < nu; i += 32) xu[nx + i] = ur[i];
          __syncwarp();
          nn_eval_warp<true>(N, xu, f, ya, yb, Ja, Jb, AB, sd, lane);
          for (int o = lane; o < nx * nin; o += 32) lAB[o] = AB[o];
          __syncwarp();
          ABk = lAB;
          if (P.lin_dare) {
            const int dit = dare_sda_warp(nx, nu, lAB, lAB + nx * nx, sQ, P.Rinv, lP, lws, lane);
            if (dit < 0) {           // no stabilising solution for this linearisation: reported as data, like the host design's error
              if (lane == 0) { P.status[p] = -20; P.iters[p] = 0; if (P.inner_iters) P.inner_iters[p] = 0; }
              p = -1;                // this warp sits the round out
            } else Wt = lP;
          }
        }
        if (p >= 0) {
          mu = 0.0;
          if (SB) { for (int r = lane; r < ms; r += 32) gy[r] = P.warm_y ? P.warm_y[p * ny + nz + r] : 0.0; }
          if (EQ) { for (int i = lane; i < nx; i += 32) syg[i] = P.warm_y ? P.warm_y[p * ny + nz + ms + i] : 0.0; }
#pragma unroll
          for (int i = 0; i < ROWS; i++) {
            const int e = lane + 32 * i;
            yd[i] = 0.0;
            if (e < nz) {
              double u0v;
              if (P.warm_u) u0v = P.warm_u[p * nz + e];
              else { const double r0 = ur[e % nu]; u0v = r0 < sLb[e] ? sLb[e] : (r0 > sUb[e] ? sUb[e] : r0); }
              su[e] = u0v;
              if (P.warm_y) yd[i] = P.warm_y[p * ny + e];
            }
          }
          __syncwarp();
        }
        status = -2; sqp_it = 0; inner_total = 0; step = 0.0; qp_rd = 0.0; Jcur = 0.0; have_traj = false;
      }
    }
    if (P.sync_rounds) {
      if (!__syncthreads_or(p >= 0)) break;
      if (p < 0) continue;
    } else if (p < 0) break;        // no warp gets a second problem: nothing to keep in step (host's choice)
    sqp_it++;
    bool fell = false;
    do {
      NMPC_PROF(4);
      // ---------------------------------------------------------------- 1. linearise along the trajectory of u
      for (int a = 0; a < nz; a++)
        for (int c = lane; c < ldk; c += 32) K[a * ldk + c] = c < nz ? sHc[a * nz + c] : 0.0;
      if (SB) { for (int o = lane; o < H * nx * ldg; o += 32) G0[o] = 0.0; }      // columns beyond the current stage must read as zero in the row products
      else { for (int o = lane; o < nx * nz; o += 32) { G0[o] = 0.0; G1[o] = 0.0; } }
      for (int i = lane; i < nx; i += 32) { const double v = x0[i]; xu[i] = v; traj[i] = v; }
      __syncwarp();
      double J0 = 0.0;
      {
        double part = 0.0;
        for (int e = lane; e < nz; e += 32) scol[e] = su[e] - ur[e % nu];
        __syncwarp();
        for (int e = lane; e < nz; e += 32) {
          double s = 0.0;
  #pragma unroll 8
        for (int j = 0; j < nz; j++) s = fma(sHc[j * nz + e], scol[j], s);
          sg[e] = s;
          part = fma(0.5 * scol[e], s, part);
        }
        __syncwarp();
        for (int i = lane; i < nx; i += 32) {
          double s = 0.0;
          for (int j = 0; j < nx; j++) s = fma(sQ[j * nx + i], x0[j] - xr[j], s);
          part = fma(x0[i] - xr[i], s, part);
        }
        J0 = part;
      }
      double* Gc = G0; double* Gn = G1;
      for (int k = 0; k < H; k++) {
        double* GW = Gc;                               // W Gamma goes to the idle buffer ...
        if (SB) { Gc = G0 + (size_t)(k > 0 ? k - 1 : 0) * nx * ldg; Gn = G0 + (size_t)k * nx * ldg; GW = G1; }   // ... or, with all Gamma_k kept, to the scratch block
        for (int i = lane; i < nu; i += 32) xu[nx + i] = su[k * nu + i];
        __syncwarp();
        NMPC_PROF(0);
        if (LIN) model_step();
        else nn_eval_warp<true>(N, xu, f, ya, yb, Ja, Jb, AB, sd, lane);
        NMPC_PROF(5);
        const int ncol = (k + 1) * nu;                 // non-zero columns of Gamma_{k+1}
        // Gamma_{k+1} = A_k Gamma_k, block k = B_k            (stored [i][c], c fastest)
        for (int i = 0; i < nx; i++)
          for (int c = lane; c < ncol; c += 32) {
            double s;
            if (c >= k * nu) s = ABk[(nx + c - k * nu) * nx + i];
            else {
              s = 0.0;
#pragma unroll 4
              for (int j = 0; j < nx; j++) s = fma(ABk[j * nx + i], Gc[j * ldg + c], s);
            }
            Gn[i * ldg + c] = s;
          }
        const double* W = (k + 1 == H) ? Wt : sQ;
        for (int i = lane; i < nx; i += 32) { const double v = f[i]; xu[i] = v; traj[(k + 1) * nx + i] = v; se[i] = v - xr[i]; }
        __syncwarp();
        for (int i = lane; i < nx; i += 32) {
          double s = 0.0;
          for (int j = 0; j < nx; j++) s = fma(W[j * nx + i], se[j], s);
          sWe[i] = s;
          J0 = fma(se[i], s, J0);
        }
        // W Gamma into the idle buffer
        for (int i = 0; i < nx; i++)
          for (int c = lane; c < ncol; c += 32) {
            double s = 0.0;
#pragma unroll 4
            for (int j = 0; j < nx; j++) s = fma(W[j * nx + i], Gn[j * ldg + c], s);
            GW[i * ldg + c] = s;
          }
        __syncwarp();
        NMPC_PROF(6);
        // K += 2 Gamma' (W Gamma),  g += 2 Gamma' (W e).  Each lane owns the columns c = lane + 32 r of K; four rows a are
        // processed per pass with all loads issued before the dependent FMAs (the rows are independent, but K, Gamma and
        // W Gamma share one shared-memory allocation, so the compiler must be told by construction).
        for (int a0 = 0; a0 < ncol; a0 += 4) {
          double acc[4][ROWS];
#pragma unroll
          for (int u = 0; u < 4; u++)
#pragma unroll
            for (int r = 0; r < ROWS; r++) acc[u][r] = 0.0;
#pragma unroll 2
          for (int i = 0; i < nx; i++) {
            double ga[4], gc[ROWS];
#pragma unroll
            for (int u = 0; u < 4; u++) ga[u] = (a0 + u < ncol) ? Gn[i * ldg + a0 + u] : 0.0;
#pragma unroll
            for (int r = 0; r < ROWS; r++) { const int c = lane + 32 * r; gc[r] = (c < ncol) ? GW[i * ldg + c] : 0.0; }
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
              for (int r = 0; r < ROWS; r++) acc[u][r] = fma(ga[u], gc[r], acc[u][r]);
          }
#pragma unroll
          for (int u = 0; u < 4; u++)
#pragma unroll
            for (int r = 0; r < ROWS; r++) {
              const int c = lane + 32 * r;
              if (a0 + u < ncol && c < ncol) K[(a0 + u) * ldk + c] = fma(2.0, acc[u][r], K[(a0 + u) * ldk + c]);
            }
        }
        for (int a = lane; a < ncol; a += 32) {
          double s = 0.0;
          for (int i = 0; i < nx; i++) s = fma(Gn[i * ldg + a], sWe[i], s);
          sg[a] = fma(2.0, s, sg[a]);
        }
        __syncwarp();
        NMPC_PROF(7);
        // the buffer that held W Gamma must again read as Gamma_k = 0 beyond its columns for the next step: it is fully
        // rewritten for c < ncol + nu next time and never read beyond, so no clearing is needed
        if (!SB) { double* tp = Gc; Gc = Gn; Gn = tp; }
      }
      J0 = warp_sum(J0);
      NMPC_PROF(0);
      // ---------------------------------------------------------------- 2. q = g - K u ; K <- (K + (sigma + rho) I)^-1
      for (int e = lane; e < nz; e += 32) {
        double s = 0.0;
#pragma unroll 8
        for (int j = 0; j < nz; j++) s = fma(K[j * ldk + e], su[j], s);
        sq[e] = sg[e] - s;
      }
      __syncwarp();
      for (int e = lane; e < nz; e += 32) K[e * ldk + e] += sig_rho;
      const double* GH = SB ? G0 + (size_t)(H - 1) * nx * ldg : Gc;      // Gamma_H
      double cviol = 0.0;                  // l1 violation of the nonlinear constraints at u: |e_H|_1 (all lanes) + state-box part (per-lane partial, reduced below)
      double cviol_sb = 0.0;
      if (EQ) {
        double n2sum = 0.0, eh2 = 0.0;
        for (int i = 0; i < nx; i++) {     // per terminal row: |G_i|^2, G_i u  (warp reductions)
          double n2 = 0.0, gu = 0.0;
          for (int cidx = lane; cidx < nz; cidx += 32) { const double gv = GH[i * ldg + cidx]; n2 = fma(gv, gv, n2); gu = fma(gv, su[cidx], gu); }
          n2 = warp_sum(n2); gu = warp_sum(gu);
          const double eh = se[i];
          cviol += fabs(eh);
          n2sum += n2; eh2 = fma(eh, eh, eh2);
          if (lane == 0) {
            const double re = P.rho_eq_scale * rho / fmax(n2, 1e-12);
            srhoe[i] = re; sb[i] = gu - eh; szg[i] = gu; sysg[i] = syg[i] / re;     // OSQP warm start: z = A x, y kept
          }
        }
        __syncwarp();
        if (P.term_ball) {                 // one common step size on the ball's rows (the projection needs a uniform metric)
          const double re = rho * (double)nx / fmax(n2sum, 1e-12);
          if (lane == 0) for (int i = 0; i < nx; i++) { srhoe[i] = re; sysg[i] = syg[i] / re; }
          cviol = fmax(sqrt(eh2) - rad, 0.0);
          __syncwarp();
        }
        for (int a = 0; a < nz; a++) {     // K += G' diag(rho_e) G
          for (int i = 0; i < nx; i++) {
            const double gai = srhoe[i] * GH[i * ldg + a];
#pragma unroll
            for (int r = 0; r < ROWS; r++) {
              const int cidx = lane + 32 * r;
              if (cidx < nz) K[a * ldk + cidx] = fma(gai, GH[i * ldg + cidx], K[a * ldk + cidx]);
            }
          }
        }
      }
      if (SB) {
        // per state-box row r = (k-1) nx + i (k = 1..H): step size, linearised bounds, OSQP warm start z = G u
        for (int r = lane; r < ms; r += 32) {
          const double* Gr = G0 + (size_t)r * ldg;
          double n2 = 0.0, gu = 0.0;
#pragma unroll 4
          for (int cidx = 0; cidx < nz; cidx++) { const double gv = Gr[cidx]; n2 = fma(gv, gv, n2); gu = fma(gv, su[cidx], gu); }
          const int i = r % nx;
          const double xk = traj[nx + r];                   // x_k[i] of the current trajectory (traj holds k = 0..H)
          const double re = rho / fmax(n2, 1e-12);
          grho[r] = re; glo[r] = sXmin[i] - xk + gu; ghi[r] = sXmax[i] - xk + gu; gz[r] = gu; gys[r] = gy[r] / re;
          cviol_sb += fmax(xk - sXmax[i], 0.0) + fmax(sXmin[i] - xk, 0.0);
        }
        __syncwarp();
        // K += G' diag(rho_g) G : each lane owns the columns c = lane + 32 r of K, four rows a per pass
        for (int a0 = 0; a0 < nz; a0 += 4) {
          double acc[4][ROWS];
#pragma unroll
          for (int u = 0; u < 4; u++)
#pragma unroll
            for (int r2 = 0; r2 < ROWS; r2++) acc[u][r2] = 0.0;
          for (int r = 0; r < ms; r++) {
            const double* Gr = G0 + (size_t)r * ldg;
            const double rg = grho[r];
            double ga[4], gc[ROWS];
#pragma unroll
            for (int u = 0; u < 4; u++) ga[u] = (a0 + u < nz) ? rg * Gr[a0 + u] : 0.0;
#pragma unroll
            for (int r2 = 0; r2 < ROWS; r2++) { const int c = lane + 32 * r2; gc[r2] = (c < nz) ? Gr[c] : 0.0; }
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
              for (int r2 = 0; r2 < ROWS; r2++) acc[u][r2] = fma(ga[u], gc[r2], acc[u][r2]);
          }
#pragma unroll
          for (int u = 0; u < 4; u++)
#pragma unroll
            for (int r2 = 0; r2 < ROWS; r2++) {
              const int c = lane + 32 * r2;
              if (a0 + u < nz && c < nz) K[(a0 + u) * ldk + c] += acc[u][r2];
            }
        }
      }
      __syncwarp();
      // In-place Gauss-Jordan inverse of the SPD matrix K (no pivoting).  Each lane owns the columns j = lane + 32 r; per
      // pivot the scaled pivot row sits in registers and the rows are swept four at a time, loads first (rows are
      // independent; see the note on aliasing above).
      for (int pv = 0; pv < nz; pv++) {
        const double dinv = 1.0 / K[pv * ldk + pv];
        for (int i = lane; i < nz; i += 32) scol[i] = K[i * ldk + pv];
        __syncwarp();
        double kp[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; r++) {
          const int j = lane + 32 * r;
          kp[r] = 0.0;
          if (j < nz) {
            kp[r] = (j == pv) ? dinv : K[pv * ldk + j] * dinv;      // with K[i][pv] read as 0 below, kp[pv] = dinv yields -f dinv
            K[pv * ldk + j] = kp[r];
          }
        }
        for (int i0 = 0; i0 < nz; i0 += 4) {
          double fct[4], kv[4][ROWS];
#pragma unroll
          for (int u = 0; u < 4; u++) { const int i = i0 + u; fct[u] = (i < nz && i != pv) ? scol[i] : 0.0; }
#pragma unroll
          for (int u = 0; u < 4; u++)
#pragma unroll
            for (int r = 0; r < ROWS; r++) {
              const int i = i0 + u, j = lane + 32 * r;
              kv[u][r] = (i < nz && j < nz && j != pv) ? K[i * ldk + j] : 0.0;
            }
#pragma unroll
          for (int u = 0; u < 4; u++)
#pragma unroll
            for (int r = 0; r < ROWS; r++) {
              const int i = i0 + u, j = lane + 32 * r;
              if (i < nz && i != pv && j < nz) K[i * ldk + j] = fma(-fct[u], kp[r], kv[u][r]);
            }
        }
        __syncwarp();
      }
      NMPC_PROF(1);
      // ---------------------------------------------------------------- 3. ADMM (box-only form of admm_onchip.cuh)
      double c[ROWS], qv[ROWS], xs[ROWS], tlast[ROWS], ylast[ROWS];
      double qn = 0.0;
#pragma unroll
      for (int i = 0; i < ROWS; i++) {
        const int e = lane + 32 * i;
        c[i] = qv[i] = xs[i] = tlast[i] = ylast[i] = 0.0;
        if (e < nz) {
          const double u0v = su[e], ys0 = yd[i] / rho;
          qv[i] = sq[e];
          xs[i] = u0v;
          c[i] = fma(oma, u0v, ys0);
          sr[e] = fma(rho, u0v - ys0, fma(sigma, u0v, -qv[i]));
          qn = fmax(qn, fabs(qv[i]));
        }
      }
      qn = warp_max(qn);
      __syncwarp();
      int it = 0;
      double rp = 0.0, rd = 0.0;
      bool qp_conv = false;
      while (true) {
        bool conv = false;
        for (int ii = 0; ii < P.check_every; ii++) {
          const bool chk = (ii == P.check_every - 1);
          if (SB) {            // right-hand side += G' (rho_g (z_g - ys_g)) over the state-box rows
            for (int r = lane; r < ms; r += 32) gm[r] = grho[r] * (gz[r] - gys[r]);
            __syncwarp();
            for (int e = lane; e < nz; e += 32) {
              double a = sr[e];
#pragma unroll 4
              for (int r = 0; r < ms; r++) a = fma(G0[(size_t)r * ldg + e], gm[r], a);
              sr[e] = a;
            }
            __syncwarp();
          }
          if (EQ) {            // right-hand side += G' (rho_e (z_g - ys_g))
            for (int i = lane; i < nx; i += 32) ssm[i] = srhoe[i] * (szg[i] - sysg[i]);
            __syncwarp();
            for (int e = lane; e < nz; e += 32) {
              double a = sr[e];
              for (int i = 0; i < nx; i++) a = fma(GH[i * ldg + e], ssm[i], a);
              sr[e] = a;
            }
            __syncwarp();
          }
          double t[ROWS];
#pragma unroll
          for (int i = 0; i < ROWS; i++) t[i] = 0.0;
          // lanes beyond nz in the last row group read the (finite) padding / next row of K and discard the result below
#pragma unroll 8
          for (int j = 0; j < nz; j++) {
            const double rj = sr[j];
#pragma unroll
            for (int i = 0; i < ROWS; i++) t[i] = fma(K[j * ldk + ((lane + 32 * i) < nz ? (lane + 32 * i) : 0)], rj, t[i]);
          }
          __syncwarp();       // every lane has consumed sr
          double nA = 0.0, nD = 0.0;
          if (chk) { rp = 0.0; rd = 0.0; }
          if (SB) {            // state-box rows: t_g = G x~ ; z_g+ = clip(alpha t_g + (1 - alpha) z_g + ys_g) ; ys_g+ = w_g - z_g+
#pragma unroll
            for (int r2 = 0; r2 < ROWS; r2++) { const int e = lane + 32 * r2; if (e < nz) sv[e] = t[r2]; }
            __syncwarp();
            for (int r = lane; r < ms; r += 32) {
              const double* Gr = G0 + (size_t)r * ldg;
              double tg = 0.0;
#pragma unroll 4
              for (int cidx = 0; cidx < nz; cidx++) tg = fma(Gr[cidx], sv[cidx], tg);
              const double wg = fma(alpha, tg, fma(oma, gz[r], gys[r]));
              const double lo = glo[r], hi = ghi[r];
              const double zg = wg < lo ? lo : (wg > hi ? hi : wg);
              const double ysn = wg - zg;
              if (chk) { rp = fmax(rp, fabs(tg - zg)); nA = fmax(nA, fmax(fabs(tg), fabs(zg))); }
              gz[r] = zg; gys[r] = ysn; gn[r] = grho[r] * (ysn - tg);
            }
            __syncwarp();
          }
          if (EQ) {            // terminal rows: t_g = G x~ ; z_g+ = b ; ys_g+ = alpha t_g + (1 - alpha) z_g + ys_g - b
            double bscale = 0.0;             // ball: z_g+ = b + bscale (w_g - b), bscale = min(1, rad / |w_g - b|_2)
            if (P.term_ball) {
              double d2 = 0.0;
              for (int i = 0; i < nx; i++) {
                double part = 0.0;
#pragma unroll
                for (int r = 0; r < ROWS; r++) { const int e = lane + 32 * r; if (e < nz) part = fma(GH[i * ldg + e], t[r], part); }
                const double zg_old = szg[i], ysg_old = sysg[i], bi = sb[i];
                const double tgi = warp_sum(part);
                const double dv = fma(alpha, tgi, fma(oma, zg_old, ysg_old)) - bi;
                d2 = fma(dv, dv, d2);
                if (lane == 0) stg[i] = tgi;
              }
              bscale = d2 > rad * rad ? rad / sqrt(d2) : 1.0;
              __syncwarp();
            }
            for (int i = 0; i < nx; i++) {
              double tgi;
              const double zg_old = szg[i], ysg_old = sysg[i], bi = sb[i], re = srhoe[i];     // read before the reduction (a convergence point) ...
              if (P.term_ball) tgi = stg[i];
              else {
                double part = 0.0;
#pragma unroll
                for (int r = 0; r < ROWS; r++) { const int e = lane + 32 * r; if (e < nz) part = fma(GH[i * ldg + e], t[r], part); }
                tgi = warp_sum(part);
              }
              const double wmb = fma(alpha, tgi, fma(oma, zg_old, ysg_old)) - bi;
              const double zgn = fma(bscale, wmb, bi);
              const double ysgn = wmb - bscale * wmb;
              if (chk) { rp = fmax(rp, fabs(tgi - zgn)); nA = fmax(nA, fmax(fabs(tgi), fabs(zgn))); }
              __syncwarp();
              if (lane == 0) { szg[i] = zgn; sysg[i] = ysgn; ssn[i] = re * (ysgn - tgi); stg[i] = tgi; }   // ... written after it
            }
            __syncwarp();
          }
#pragma unroll
          for (int i = 0; i < ROWS; i++) {
            const int e = lane + 32 * i;
            if (e < nz) {
              const double wv = fma(alpha, t[i], c[i]);
              const double lo = sLb[e], hi = sUb[e];
              const double zn = wv < lo ? lo : (wv > hi ? hi : wv);
              if (chk) {
                double pc = fma(-sig_rho, t[i], sr[e]);
                if (EQ) for (int k2 = 0; k2 < nx; k2++) pc = fma(GH[k2 * ldg + e], ssn[k2], pc);     // Kgn x~ + G' y_g from the cached factor
                if (SB) {
#pragma unroll 4
                  for (int r = 0; r < ms; r++) pc = fma(G0[(size_t)r * ldg + e], gn[r], pc);
                }
                const double ybv = rho * (wv - zn);
                rp = fmax(rp, fabs(t[i] - zn));
                rd = fmax(rd, fabs(pc + qv[i] + ybv));
                nA = fmax(nA, fmax(fabs(t[i]), fabs(zn)));
                nD = fmax(nD, fmax(fabs(pc), fabs(ybv)));
                tlast[i] = t[i]; ylast[i] = ybv;
              }
              c[i] = fma(-alpha, zn, wv);
              xs[i] = fma(alpha, t[i], oma * xs[i]);
              sr[e] = fma(rho, fma(2.0, zn, -wv), fma(sigma, xs[i], -qv[i]));
            }
          }
          __syncwarp();
          if (chk) {
            rp = warp_max(rp); rd = warp_max(rd); nA = warp_max(nA); nD = warp_max(nD);
            conv = (rp <= P.eps_abs + P.eps_rel * nA) && (rd <= P.eps_abs + P.eps_rel * fmax(nD, qn));
          }
        }
        it += P.check_every;
        if (conv) { qp_conv = true; break; }
        if (it >= max_iter) break;
      }
      inner_total += it;
      NMPC_PROF(2);
      qp_rd = rd;
      bool qp_failed = false;
      if (SB) {
        double ymax = 0.0;
        for (int r = lane; r < ms; r += 32) { const double yv = grho[r] * gys[r]; gy[r] = yv; ymax = fmax(ymax, fabs(yv)); }
        mu = fmax(mu, 1.1 * warp_max(ymax));
        cviol += warp_sum(cviol_sb);
        qp_failed = !qp_conv;
        __syncwarp();
      }
      if (EQ) {
        double ymax = 0.0;
        double y2 = 0.0;
        for (int i = 0; i < nx; i++) { const double yv = srhoe[i] * sysg[i]; ymax = fmax(ymax, fabs(yv)); y2 = fma(yv, yv, y2); if (lane == 0) syg[i] = yv; }
        mu = fmax(mu, 1.1 * (P.term_ball ? sqrt(y2) : ymax));      // exact-penalty weight: dual norm of the violation measure (l1 -> max, l2 -> l2)
        qp_failed = qp_failed || !qp_conv;          // linearised rows + input box not solvable within the inner cap
        __syncwarp();
      }
      // ---------------------------------------------------------------- 4. step, acceptance, line search
      double dmax = 0.0, gd = 0.0;
#pragma unroll
      for (int i = 0; i < ROWS; i++) {
        const int e = lane + 32 * i;
        if (e < nz) {
          const double d = tlast[i] - su[e];
          dmax = fmax(dmax, fabs(d));
          gd = fma(sg[e], d, gd);
          sv[e] = tlast[i];
          yd[i] = ylast[i];
        }
      }
      dmax = warp_max(dmax); gd = warp_sum(gd);
      step = dmax;
      __syncwarp();
      if (LIN) {                         // the QP is the problem: its solution is the answer, its convergence the status
        for (int e = lane; e < nz; e += 32) su[e] = sv[e];
        __syncwarp();
        status = qp_conv ? 1 : -2; step = rp; have_traj = false;
        break;
      }
      if (EQ || SB) {
        if (qp_failed) { status = -3; have_traj = false; break; }
        gd -= mu * cviol;                // directional derivative of the l1 merit (the full step zeroes the linearised rows)
        J0 += mu * cviol;
      }
      if (dmax <= P.sqp_tol) {           // converged: take the full step
        for (int e = lane; e < nz; e += 32) su[e] = sv[e];
        __syncwarp();
        status = 1; have_traj = false;
        break;
      }
      bool ok = false;
      double tls = 1.0;
      for (int ls = 0; ls <= P.ls_max; ls++) {
#pragma unroll
        for (int i = 0; i < ROWS; i++) {
          const int e = lane + 32 * i;
          if (e < nz) sv[e] = fma(tls, tlast[i] - su[e], su[e]);
        }
        __syncwarp();
        const double Jc = rollout_cost(sv);
        double merit = Jc;
        if (EQ || SB) {
          double cv = 0.0;
          if (EQ) {
            double c1 = 0.0, c2 = 0.0;
            for (int i = 0; i < nx; i++) { const double ev = traj[H * nx + i] - xr[i]; c1 += fabs(ev); c2 = fma(ev, ev, c2); }
            cv += P.term_ball ? fmax(sqrt(c2) - rad, 0.0) : c1;
          }
          if (SB) {
            double part = 0.0;
            for (int r = lane; r < ms; r += 32) { const double xk = traj[nx + r]; const int i = r % nx; part += fmax(xk - sXmax[i], 0.0) + fmax(sXmin[i] - xk, 0.0); }
            cv += warp_sum(part);
          }
          merit = fma(mu, cv, Jc);
        }
        if (merit <= J0 + P.ls_c1 * tls * gd + P.ls_noise * fmax(1.0, fabs(J0))) { ok = true; Jcur = Jc; break; }
        tls *= 0.5;
      }
      if (!ok) { status = 2; have_traj = false; break; }       // stalled at a kink: keep u
      for (int e = lane; e < nz; e += 32) su[e] = sv[e];
      __syncwarp();
      have_traj = true;
      fell = true;
    } while (0);
    if (fell && sqp_it < P.sqp_max_iter) continue;       // next round: another SQP iteration of the same problem
    NMPC_PROF(3);
    if (!have_traj) Jcur = rollout_cost(su);
    // ------------------------------------------------------------------ outputs
    for (int e = lane; e < nz; e += 32) {
      const double uv = su[e], ev = uv - ur[e % nu];
      if (P.u) P.u[p * nz + e] = uv;
      if (P.e_u) P.e_u[p * nz + e] = ev;
      if (P.u0 && e < nu) P.u0[p * nu + e] = uv;
    }
    if (P.y) {
#pragma unroll
      for (int i = 0; i < ROWS; i++) { const int e = lane + 32 * i; if (e < nz) P.y[p * ny + e] = yd[i]; }
      if (SB) for (int r = lane; r < ms; r += 32) P.y[p * ny + nz + r] = gy[r];
      if (EQ) for (int i = lane; i < nx; i += 32) P.y[p * ny + nz + ms + i] = syg[i];
    }
    for (int o = lane; o < (H + 1) * nx; o += 32) {
      const double xv = traj[o];
      if (P.x) P.x[p * (H + 1) * nx + o] = xv;
      if (P.e_x) P.e_x[p * (H + 1) * nx + o] = xv - xr[o % nx];
    }
    if (lane == 0) {
      if (P.objective) P.objective[p] = Jcur;
      P.status[p] = status; P.iters[p] = sqp_it;
      if (P.inner_iters) P.inner_iters[p] = inner_total;
      if (P.step) P.step[p] = step;
      if (P.qp_dres) P.qp_dres[p] = qp_rd;
    }
    __syncwarp();
    p = -1;
  }
}

}  // namespace mpcb
