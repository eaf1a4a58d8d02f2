// C ABI of the nonlinear path (include/mpcb200.h, section "Nonlinear path"): network upload, batched rollout /
// Jacobian entry points, NMPC controller design (linearisation at the design reference on the GPU, DARE + rho on the
// host) and the batched SQP solve.  No CPU fallback: every compute entry needs an sm_100 device.
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mpcb200.h"
#include "api_common.hpp"
#include "host_design.hpp"
#include "nmpc.cuh"
#include "nmpc_launch.hpp"
#include "multi_device.hpp"

using mpcb::api_fail;
using mpcb::DevBuf;
using mpcb::PinBuf;

struct mpcb_nn {
  int device = 0;
  int sm_count = 0;
  mpcb::NetDev net{};
  DevBuf<double> weights;
  cudaStream_t stream = nullptr;
  DevBuf<double> in0, in1, out0, out1, out2;   // staging for the host-pointer entry points
  size_t smem_set[2] = {0, 0};
};

struct mpcb_nmpc {
  mpcb_nn* nn = nullptr;
  mpcb_nmpc_settings st{};
  int H = 0, nz = 0, rows = 0, warps = 0;
  int warps_lin = 0;          // CTA width of the re-linearised (LIN) kernels: a little more shared memory per warp
  bool terminal_eq = false, terminal_ball = false, state_box = false;   // terminal_eq: the kernels with terminal rows (equality, or the contractive ball)
  double rho = 0.0;
  mpcb::Mat A, B, P;
  DevBuf<double> Q, Pt, Hc, lb, ub, xmin, xmax, Rinv;
  DevBuf<unsigned long long> counter;
  // host-entry workspaces
  DevBuf<double> x0, xref, uref, warm_u, warm_y, u, e_u, x, e_x, u0, obj, y, step, dres;
  DevBuf<int32_t> status, iters, inner;
  PinBuf<int32_t> stage_int;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  mpcb_timing timing{};
  size_t smem_set = 0, smem_set_lin = 0;
  std::vector<mpcb_nmpc*> peers;   // multi-device handle (settings.qp.n_devices > 1): the controllers on the other devices
};

namespace {

int check_device(int device, int* sm_count) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return api_fail(MPCB_ERR_NO_DEVICE, "no CUDA device visible: libmpcb200 has no CPU fallback");
  }
  if (device < 0 || device >= ndev) return api_fail(MPCB_ERR_INVALID, "device out of range");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return api_fail(MPCB_ERR_NO_DEVICE, std::string("device ") + prop.name + " is not sm_100: this library is built for B200 only");
  *sm_count = prop.multiProcessorCount;
  return MPCB_OK;
}

int validate_nn(const mpcb_nn_desc* d) {
  if (!d) return api_fail(MPCB_ERR_INVALID, "null network description");
  if (d->arch < MPCB_NN_FNN || d->arch > MPCB_NN_DENSENET)
    return api_fail(MPCB_ERR_INVALID, "unknown network architecture (fnn, resnet, polynet and densenet are supported)");
  if (d->activation < MPCB_ACT_RELU || d->activation > MPCB_ACT_IDENTITY) return api_fail(MPCB_ERR_INVALID, "unknown activation id");
  if (d->nx <= 0 || d->nu <= 0 || d->n_neurons <= 0 || d->n_hidden < 0) return api_fail(MPCB_ERR_INVALID, "bad network sizes");
  if (!d->W_in || !d->W_out || (d->n_hidden > 0 && (!d->W_hidden || !d->b_hidden))) return api_fail(MPCB_ERR_INVALID, "null weight pointer");
  return MPCB_OK;
}

template <bool JAC>
int launch_nn_batch(mpcb_nn* n, const mpcb::NnBatchParams& P, cudaStream_t st) {
  const size_t smem = mpcb::nn_batch_smem_bytes(n->net, JAC);
  if (smem > 200 * 1024) return api_fail(MPCB_ERR_INVALID, "network too large for the shared-memory resident kernels");
  auto kern = mpcb::nn_batch_kernel<JAC>;
  if (smem > 48 * 1024 && smem > n->smem_set[JAC ? 1 : 0]) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    n->smem_set[JAC ? 1 : 0] = smem;
  }
  const long long blocks = (P.batch + mpcb::NN_WARPS - 1) / mpcb::NN_WARPS;
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(blocks, (long long)n->sm_count * 8));
  kern<<<grid, mpcb::NN_THREADS, smem, st>>>(P);
  CUDA_TRY(cudaGetLastError());
  return MPCB_OK;
}

int jacobian_device(mpcb_nn* n, int64_t batch, const double* x, const double* u, double* f, double* A, double* B, cudaStream_t st) {
  mpcb::NnBatchParams P{};
  P.net = n->net; P.batch = batch; P.H = 1; P.x0 = x; P.u = u; P.x = f; P.A = A; P.B = B;
  return launch_nn_batch<true>(n, P, st);
}

// Plant step of the GPU-resident closed loop on a neural model: one warp per plant applies u0 to the network itself
// (x+ = f(x, u0): the plant IS the model, as in the reference's closed-loop test pattern), appends to the trajectories,
// accumulates the counters and prepares the next solve's warm start: inputs and input-box duals shifted by one stage
// (the last stage repeated), the duals of the state / terminal rows carried over as they are.
struct NnPlantParams {
  mpcb::NetDev net;
  int H, nz, ny, t, steps;
  long long batch;
  const double* u;          // [batch][nz] solution of this step
  const double* y;          // [batch][ny] duals of this step (null without warm start)
  const int32_t *status, *inner;
  const double* x_in;       // [batch][nx]
  double* x_out;
  double *warm_u, *warm_y;  // next step's warm start (null without warm start)
  double *x_traj, *u_traj;
  int32_t *iters_total, *unsolved;
};

__global__ void __launch_bounds__(mpcb::NN_THREADS) nn_plant_step_kernel(const NnPlantParams P) {
  extern __shared__ __align__(16) double sm[];
  const mpcb::NetSm N = mpcb::stage_network(P.net, sm, threadIdx.x, mpcb::NN_THREADS);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nx = N.nx, nu = N.nu, nin = N.nin, nn = N.nn, nz = P.nz;
  double* w = sm + P.net.weight_count() + warp * mpcb::nn_eval_scratch_doubles(P.net, false);
  double* xu = w; double* f = xu + nin; double* ya = f + nx; double* yb = ya + nn; double* sd = yb + nn;
  for (long long p = (long long)blockIdx.x * mpcb::NN_WARPS + warp; p < P.batch; p += (long long)gridDim.x * mpcb::NN_WARPS) {
    for (int i = lane; i < nx; i += 32) {
      const double v = P.x_in[p * nx + i];
      xu[i] = v;
      if (P.t == 0 && P.x_traj) P.x_traj[(p * (P.steps + 1)) * nx + i] = v;
    }
    for (int i = lane; i < nu; i += 32) {
      const double v = P.u[p * nz + i];
      xu[nx + i] = v;
      if (P.u_traj) P.u_traj[(p * P.steps + P.t) * nu + i] = v;
    }
    __syncwarp();
    mpcb::nn_eval_warp<false>(N, xu, f, ya, yb, nullptr, nullptr, nullptr, sd, lane);
    for (int i = lane; i < nx; i += 32) {
      const double v = f[i];
      P.x_out[p * nx + i] = v;
      if (P.x_traj) P.x_traj[(p * (P.steps + 1) + P.t + 1) * nx + i] = v;
    }
    if (P.warm_u)
      for (int e = lane; e < nz; e += 32) P.warm_u[p * nz + e] = P.u[p * nz + (e + nu < nz ? e + nu : e)];
    if (P.warm_y && P.y)
      for (int e = lane; e < P.ny; e += 32) P.warm_y[p * P.ny + e] = P.y[p * P.ny + ((e < nz && e + nu < nz) ? e + nu : e)];
    if (lane == 0) {
      if (P.iters_total) P.iters_total[p] = (P.t == 0 ? 0 : P.iters_total[p]) + P.inner[p];
      if (P.unsolved) P.unsolved[p] = (P.t == 0 ? 0 : P.unsolved[p]) + (P.status[p] == 1 ? 0 : 1);
    }
    __syncwarp();
  }
}

int enqueue_nmpc(mpcb_nmpc* h, const mpcb_batch_io& io, cudaStream_t st, bool lin = false) {
  mpcb_nn* n = h->nn;
  const long long Bn = io.batch;
  if (Bn <= 0) return api_fail(MPCB_ERR_INVALID, "batch must be positive");
  if (!io.x0 || !io.xref || !io.uref) return api_fail(MPCB_ERR_INVALID, "x0, xref, uref are required");
  int32_t* d_status = io.status; int32_t* d_iters = io.iters;
  if (!d_status) { CUDA_TRY(h->status.ensure(Bn)); d_status = h->status.p; }
  if (!d_iters) { CUDA_TRY(h->iters.ensure(Bn)); d_iters = h->iters.p; }
  CUDA_TRY(h->counter.ensure(1));
  CUDA_TRY(cudaMemsetAsync(h->counter.p, 0, sizeof(unsigned long long), st));
  mpcb::NmpcParams P{};
  P.net = n->net; P.Q = h->Q.p; P.Pt = h->Pt.p; P.Hc = h->Hc.p; P.lb = h->lb.p; P.ub = h->ub.p; P.H = h->H; P.nz = h->nz;
  const mpcb_settings& q = h->st.qp;
  P.rho = h->rho; P.sigma = q.sigma; P.alpha = q.alpha; P.eps_abs = q.eps_abs; P.eps_rel = q.eps_rel; P.max_iter = q.max_iter; P.check_every = q.check_every;
  P.sqp_tol = h->st.sqp_tol; P.ls_c1 = h->st.ls_armijo; P.ls_noise = h->st.ls_noise; P.rho_eq_scale = q.rho_eq_scale; P.term_ball = h->terminal_ball ? 1 : 0; P.xmin = h->xmin.p; P.xmax = h->xmax.p; P.sqp_max_iter = h->st.sqp_max_iter; P.ls_max = h->st.ls_max_halvings;
  P.batch = Bn; P.x0 = io.x0; P.xref = io.xref; P.uref = io.uref; P.xref_bc = io.xref_broadcast; P.uref_bc = io.uref_broadcast;
  P.warm_u = io.warm_u; P.warm_y = io.warm_y;
  P.u = io.u; P.e_u = io.e_u; P.x = io.x; P.e_x = io.e_x; P.u0 = io.u0; P.objective = io.objective; P.y = io.y;
  P.status = d_status; P.iters = d_iters; P.inner_iters = io.inner_iters; P.step = io.prim_res; P.qp_dres = io.dual_res;
  P.counter = h->counter.p;
  P.Rinv = h->Rinv.p; P.lin_dare = 1;
  if (lin && h->warps_lin == 0) return api_fail(MPCB_ERR_INVALID, "problem too large for the re-linearised kernel");
  if (lin && !h->Rinv.p) return api_fail(MPCB_ERR_NUMERIC, "re-linearised solve: the per-problem Riccati equation needs a non-singular R and nu <= 3 nx");
  const int warps = lin ? h->warps_lin : h->warps;
  const int threads = warps * 32;
  const size_t smem = mpcb::nmpc_smem_bytes(n->net, h->H, h->nz, warps, h->state_box, lin);
  const long long blocks = (Bn + warps - 1) / warps;
  const int per_sm = std::max<int>(1, (int)((226 * 1024) / smem));
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(blocks, (long long)n->sm_count * per_sm));
  P.sync_rounds = Bn > (long long)grid * warps ? 1 : 0;      // off only when no warp gets a second problem.  Measured: +33 % at 49 problems per warp;
                                                              // at 3 per warp +10..18 % for the 6-9-iteration networks, -4 % for the 2-iteration ResNet
  cudaError_t e;
  if (h->rows < 1 || h->rows > 4) return api_fail(MPCB_ERR_INVALID, "NMPC supports nu*horizon <= 128");
  if (lin) e = mpcb::launch_lin(h->terminal_eq, h->state_box, h->rows, P, grid, threads, smem, &h->smem_set_lin, st);
  else e = mpcb::launch_sqp(h->terminal_eq, h->state_box, h->rows, P, grid, threads, smem, &h->smem_set, st);
  if (e != cudaSuccess) return api_fail(MPCB_ERR_CUDA, std::string("nmpc_sqp_kernel launch: ") + cudaGetErrorString(e));
  h->timing.kernel_launches = 1;
  h->timing.batch = Bn;
  return MPCB_OK;
}

}  // namespace

#ifdef MPCB_NMPC_PROF
extern "C" int mpcb_debug_nmpc_prof(unsigned long long* out8, int reset) {   // development builds only (not in the header)
  cudaDeviceSynchronize();
  if (out8) cudaMemcpyFromSymbol(out8, mpcb::g_nmpc_prof, 8 * sizeof(unsigned long long));
  if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(mpcb::g_nmpc_prof, z, sizeof(z)); }
  return 0;
}
#endif

extern "C" {

int mpcb_create_nn(const mpcb_nn_desc* d, int32_t device, mpcb_nn** out) {
  if (!out) return api_fail(MPCB_ERR_INVALID, "mpcb_create_nn: null output");
  *out = nullptr;
  int rc = validate_nn(d);
  if (rc != MPCB_OK) return rc;
  int sms = 0;
  rc = check_device(device, &sms);
  if (rc != MPCB_OK) return rc;
  mpcb_nn* n = new mpcb_nn();
  n->device = device; n->sm_count = sms;
  mpcb::NetDev& N = n->net;
  N.arch = d->arch; N.act = d->activation; N.nx = d->nx; N.nu = d->nu; N.nn = d->n_neurons; N.nh = d->n_hidden; N.nin = d->nx + d->nu;
  const size_t n1 = (size_t)N.nn * N.nin, n2 = N.wh_count(), n3 = (size_t)N.nh * N.nn, n4 = N.wout_count();
  std::vector<double> w(n1 + n2 + n3 + n4);
  std::memcpy(w.data(), d->W_in, n1 * sizeof(double));
  if (n2) std::memcpy(w.data() + n1, d->W_hidden, n2 * sizeof(double));
  if (n3) std::memcpy(w.data() + n1 + n2, d->b_hidden, n3 * sizeof(double));
  std::memcpy(w.data() + n1 + n2 + n3, d->W_out, n4 * sizeof(double));
  for (double v : w)
    if (!std::isfinite(v)) { delete n; return api_fail(MPCB_ERR_INVALID, "network weights contain non-finite values"); }
  if (mpcb::upload(n->weights, w.data(), w.size()) != cudaSuccess) { cudaGetLastError(); delete n; return api_fail(MPCB_ERR_CUDA, "weight upload failed"); }
  N.W_in = n->weights.p; N.W_h = N.W_in + n1; N.b_h = N.W_h + n2; N.W_out = N.b_h + n3;
  if (cudaStreamCreateWithFlags(&n->stream, cudaStreamNonBlocking) != cudaSuccess) { n->weights.release(); delete n; return api_fail(MPCB_ERR_CUDA, "cudaStreamCreate failed"); }
  *out = n;
  return MPCB_OK;
}

void mpcb_destroy_nn(mpcb_nn* n) {
  if (!n) return;
  cudaSetDevice(n->device);
  if (n->stream) { cudaStreamSynchronize(n->stream); cudaStreamDestroy(n->stream); }
  n->weights.release(); n->in0.release(); n->in1.release(); n->out0.release(); n->out1.release(); n->out2.release();
  delete n;
}

int mpcb_nn_rollout_batch_device(mpcb_nn* n, int64_t batch, int32_t horizon, const double* x0, const double* u, double* x, void* cuda_stream) {
  if (!n || !x0 || !u || !x) return api_fail(MPCB_ERR_INVALID, "null argument");
  if (batch <= 0 || horizon <= 0) return api_fail(MPCB_ERR_INVALID, "batch and horizon must be positive");
  CUDA_TRY(cudaSetDevice(n->device));
  mpcb::NnBatchParams P{};
  P.net = n->net; P.batch = batch; P.H = horizon; P.x0 = x0; P.u = u; P.x = x;
  return launch_nn_batch<false>(n, P, (cudaStream_t)cuda_stream);
}

int mpcb_nn_rollout_batch(mpcb_nn* n, int64_t batch, int32_t horizon, const double* x0, const double* u, double* x) {
  if (!n || !x0 || !u || !x) return api_fail(MPCB_ERR_INVALID, "null argument");
  if (batch <= 0 || horizon <= 0) return api_fail(MPCB_ERR_INVALID, "batch and horizon must be positive");
  CUDA_TRY(cudaSetDevice(n->device));
  const size_t nx = n->net.nx, nu = n->net.nu, B = (size_t)batch, H = (size_t)horizon;
  CUDA_TRY(n->in0.ensure(nx * B)); CUDA_TRY(n->in1.ensure(nu * H * B)); CUDA_TRY(n->out0.ensure(nx * (H + 1) * B));
  CUDA_TRY(cudaMemcpyAsync(n->in0.p, x0, nx * B * sizeof(double), cudaMemcpyHostToDevice, n->stream));
  CUDA_TRY(cudaMemcpyAsync(n->in1.p, u, nu * H * B * sizeof(double), cudaMemcpyHostToDevice, n->stream));
  int rc = mpcb_nn_rollout_batch_device(n, batch, horizon, n->in0.p, n->in1.p, n->out0.p, n->stream);
  if (rc != MPCB_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(x, n->out0.p, nx * (H + 1) * B * sizeof(double), cudaMemcpyDeviceToHost, n->stream));
  CUDA_TRY(cudaStreamSynchronize(n->stream));
  return MPCB_OK;
}

int mpcb_nn_jacobian_batch_device(mpcb_nn* n, int64_t batch, const double* x, const double* u, double* f, double* A, double* B, void* cuda_stream) {
  if (!n || !x || !u || !A || !B) return api_fail(MPCB_ERR_INVALID, "null argument");
  if (batch <= 0) return api_fail(MPCB_ERR_INVALID, "batch must be positive");
  CUDA_TRY(cudaSetDevice(n->device));
  return jacobian_device(n, batch, x, u, f, A, B, (cudaStream_t)cuda_stream);
}

int mpcb_nn_jacobian_batch(mpcb_nn* n, int64_t batch, const double* x, const double* u, double* f, double* A, double* B) {
  if (!n || !x || !u || !A || !B) return api_fail(MPCB_ERR_INVALID, "null argument");
  if (batch <= 0) return api_fail(MPCB_ERR_INVALID, "batch must be positive");
  CUDA_TRY(cudaSetDevice(n->device));
  const size_t nx = n->net.nx, nu = n->net.nu, Bn = (size_t)batch;
  CUDA_TRY(n->in0.ensure(nx * Bn)); CUDA_TRY(n->in1.ensure(nu * Bn));
  CUDA_TRY(n->out0.ensure(nx * Bn)); CUDA_TRY(n->out1.ensure(nx * nx * Bn)); CUDA_TRY(n->out2.ensure(nx * nu * Bn));
  CUDA_TRY(cudaMemcpyAsync(n->in0.p, x, nx * Bn * sizeof(double), cudaMemcpyHostToDevice, n->stream));
  CUDA_TRY(cudaMemcpyAsync(n->in1.p, u, nu * Bn * sizeof(double), cudaMemcpyHostToDevice, n->stream));
  int rc = jacobian_device(n, batch, n->in0.p, n->in1.p, n->out0.p, n->out1.p, n->out2.p, n->stream);
  if (rc != MPCB_OK) return rc;
  if (f) CUDA_TRY(cudaMemcpyAsync(f, n->out0.p, nx * Bn * sizeof(double), cudaMemcpyDeviceToHost, n->stream));
  CUDA_TRY(cudaMemcpyAsync(A, n->out1.p, nx * nx * Bn * sizeof(double), cudaMemcpyDeviceToHost, n->stream));
  CUDA_TRY(cudaMemcpyAsync(B, n->out2.p, nx * nu * Bn * sizeof(double), cudaMemcpyDeviceToHost, n->stream));
  CUDA_TRY(cudaStreamSynchronize(n->stream));
  return MPCB_OK;
}

// Batched Riccati equations (dare.cuh): A, B, P, status are DEVICE arrays, Q and R small HOST matrices.
int mpcb_dare_batch_device(int32_t device, int64_t batch, int32_t nx, int32_t nu, const double* A, const double* B, const double* Q, const double* R, double* P,
                           int32_t* status, void* cuda_stream) {
  if (batch <= 0 || nx <= 0 || nu <= 0 || !A || !B || !Q || !R || !P) return api_fail(MPCB_ERR_INVALID, "mpcb_dare_batch: bad arguments");
  if (nu > 3 * nx) return api_fail(MPCB_ERR_INVALID, "mpcb_dare_batch: nu <= 3 nx is required");
  int sms = 0;
  int rc = check_device(device, &sms);
  if (rc != MPCB_OK) return rc;
  const size_t smem = mpcb::dare_batch_smem_bytes(nx, nu);
  if (smem > 200 * 1024) return api_fail(MPCB_ERR_INVALID, "mpcb_dare_batch: system too large for the shared-memory resident kernel");
  mpcb::Mat Ri = mpcb::Mat::eye(nu);
  if (!mpcb::lu_solve(mpcb::Mat::from(R, nu, nu), Ri)) return api_fail(MPCB_ERR_NUMERIC, "dare: R is singular");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  double* consts = nullptr;
  CUDA_TRY(cudaMallocAsync((void**)&consts, sizeof(double) * ((size_t)nx * nx + (size_t)nu * nu), st));
  cudaError_t e = cudaMemcpyAsync(consts, Q, sizeof(double) * nx * nx, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(consts + (size_t)nx * nx, Ri.a.data(), sizeof(double) * nu * nu, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && smem > 48 * 1024) e = cudaFuncSetAttribute(mpcb::dare_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) {
    mpcb::DareBatchParams D{};
    D.nx = nx; D.nu = nu; D.batch = batch; D.A = A; D.B = B; D.Q = consts; D.Rinv = consts + (size_t)nx * nx; D.P = P; D.status = status;
    const long long blocks = (batch + mpcb::DARE_WARPS - 1) / mpcb::DARE_WARPS;
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(blocks, (long long)sms * 8));
    mpcb::dare_batch_kernel<<<grid, mpcb::DARE_WARPS * 32, smem, st>>>(D);
    e = cudaGetLastError();
  }
  cudaFreeAsync(consts, st);
  if (e != cudaSuccess) { cudaGetLastError(); return api_fail(MPCB_ERR_CUDA, std::string("mpcb_dare_batch: ") + cudaGetErrorString(e)); }
  return MPCB_OK;
}

int mpcb_dare_batch(int32_t device, int64_t batch, int32_t nx, int32_t nu, const double* A, const double* B, const double* Q, const double* R, double* P,
                    int32_t* status) {
  if (batch <= 0 || nx <= 0 || nu <= 0 || !A || !B || !Q || !R || !P) return api_fail(MPCB_ERR_INVALID, "mpcb_dare_batch: bad arguments");
  int sms = 0;
  int rc = check_device(device, &sms);
  if (rc != MPCB_OK) return rc;
  const size_t nA = (size_t)batch * nx * nx, nB = (size_t)batch * nx * nu;
  DevBuf<double> dA, dB, dP;
  DevBuf<int32_t> dS;
  auto done = [&](int code) { dA.release(); dB.release(); dP.release(); dS.release(); return code; };
  if (mpcb::upload(dA, A, nA) != cudaSuccess || mpcb::upload(dB, B, nB) != cudaSuccess || dP.ensure(nA) != cudaSuccess || dS.ensure((size_t)batch) != cudaSuccess) {
    cudaGetLastError();
    return done(api_fail(MPCB_ERR_CUDA, "mpcb_dare_batch: device buffers"));
  }
  rc = mpcb_dare_batch_device(device, batch, nx, nu, dA.p, dB.p, Q, R, dP.p, dS.p, nullptr);
  if (rc != MPCB_OK) return done(rc);
  cudaError_t e = cudaMemcpy(P, dP.p, nA * sizeof(double), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && status) e = cudaMemcpy(status, dS.p, (size_t)batch * sizeof(int32_t), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { cudaGetLastError(); return done(api_fail(MPCB_ERR_CUDA, std::string("mpcb_dare_batch: ") + cudaGetErrorString(e))); }
  return done(MPCB_OK);
}

void mpcb_default_nmpc_settings(mpcb_nmpc_settings* s) {
  if (!s) return;
  std::memset(s, 0, sizeof(*s));
  mpcb_default_settings(&s->qp);
  s->qp.eps_abs = 1e-9; s->qp.eps_rel = 0.0; s->qp.check_every = 5; s->qp.sigma = 0.0;
  s->qp.max_iter = 1000;     // inner cap per QP: smooth networks need < 200; at relu kinks the Gauss-Newton QP can be nearly singular and a 4000 cap
                             // is what a batch then waits for (reference's relu FNN fixture, 4096 problems: 69 -> 26 ms, 0.4 % fewer 'solved')
  s->sqp_tol = 1e-6; s->ls_armijo = 1e-4; s->ls_noise = 1e-10; s->sqp_max_iter = 20; s->ls_max_halvings = 12;
}

int mpcb_create_nmpc(const mpcb_nmpc_desc* d, const mpcb_nmpc_settings* settings, mpcb_nmpc** out) {
  if (!d || !out) return api_fail(MPCB_ERR_INVALID, "mpcb_create_nmpc: null argument");
  *out = nullptr;
  mpcb_nmpc_settings st;
  if (settings) st = *settings; else mpcb_default_nmpc_settings(&st);
  const mpcb_settings& q = st.qp;
  if (q.check_every <= 0 || q.max_iter <= 0 || !(q.alpha > 0 && q.alpha < 2) || !(q.sigma >= 0) || !(q.eps_abs >= 0) || !(q.eps_rel >= 0) ||
      st.sqp_max_iter <= 0 || st.ls_max_halvings < 0 || !(st.sqp_tol >= 0) || !(st.ls_armijo > 0 && st.ls_armijo < 1) || !(st.ls_noise >= 0))
    return api_fail(MPCB_ERR_INVALID, "mpcb_create_nmpc: invalid settings");
  if (d->terminal_mode != MPCB_TERMINAL_NONE && d->terminal_mode != MPCB_TERMINAL_EQUALITY && d->terminal_mode != MPCB_TERMINAL_CONTRACTIVE)
    return api_fail(MPCB_ERR_INVALID, "mpcb_create_nmpc: terminal ingredient must be 'none', 'equality' or 'contractive'");
  if (d->horizon <= 0 || !d->Q || !d->R || !d->umin || !d->umax || !d->xref || !d->uref) return api_fail(MPCB_ERR_INVALID, "mpcb_create_nmpc: bad arguments");
  std::vector<int> dev_ids;
  {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) { cudaGetLastError(); ndev = 0; }
    int rcd = mpcb::parse_devices(q, ndev, dev_ids);
    if (rcd != MPCB_OK) return rcd;
    if (!dev_ids.empty()) st.qp.device = dev_ids[0];
  }
  mpcb_nn* n = nullptr;
  int rc = mpcb_create_nn(d->nn, q.device, &n);
  if (rc != MPCB_OK) return rc;
  const int nx = n->net.nx, nu = n->net.nu, H = d->horizon, nz = nu * H;
  mpcb_nmpc* h = new mpcb_nmpc();
  h->nn = n; h->st = st; h->H = H; h->nz = nz; h->rows = (nz + 31) / 32; h->terminal_eq = d->terminal_mode != MPCB_TERMINAL_NONE; h->terminal_ball = d->terminal_mode == MPCB_TERMINAL_CONTRACTIVE; h->state_box = d->state_constraint != 0;
  if (h->state_box && (!d->xmin || !d->xmax)) { mpcb_destroy_nmpc(h); return api_fail(MPCB_ERR_INVALID, "state_constraint needs xmin and xmax"); }
  auto bail = [&](int code, const std::string& msg) { mpcb_destroy_nmpc(h); return api_fail(code, msg); };
  if (h->rows > 4) return bail(MPCB_ERR_INVALID, "NMPC supports nu*horizon <= 128");
  // CTA width: the kernel is latency bound and shared-memory limited (one K matrix per warp), so pick the width that puts
  // the most warps on an SM (227 KB); ties go to the narrower CTA.
  h->warps = 0;
  int best = 0;
  for (int w = 1; w <= mpcb::NMPC_MAX_WARPS; w++) {
    const size_t sm = mpcb::nmpc_smem_bytes(n->net, H, nz, w, h->state_box);
    if (sm > 226 * 1024) break;
    const int resident = (int)((226 * 1024) / sm) * w;
    if (resident > best) { best = resident; h->warps = w; }
  }
  if (h->warps == 0) return bail(MPCB_ERR_INVALID, "NMPC problem too large for the shared-memory resident SQP kernel");
  best = 0;
  for (int w = 1; w <= mpcb::NMPC_MAX_WARPS; w++) {
    const size_t sm = mpcb::nmpc_smem_bytes(n->net, H, nz, w, h->state_box, true);
    if (sm > 226 * 1024) break;
    const int resident = (int)((226 * 1024) / sm) * w;
    if (resident > best) { best = resident; h->warps_lin = w; }
  }

  // linearise at the design reference ON THE GPU (the reference: proceed_system_linearization, design_mpc.jl:319-323)
  {
    DevBuf<double> dx, du, dA, dB;
    if (mpcb::upload(dx, d->xref, nx) != cudaSuccess || mpcb::upload(du, d->uref, nu) != cudaSuccess || dA.ensure((size_t)nx * nx) != cudaSuccess ||
        dB.ensure((size_t)nx * nu) != cudaSuccess) { cudaGetLastError(); return bail(MPCB_ERR_CUDA, "linearisation buffers"); }
    rc = jacobian_device(n, 1, dx.p, du.p, nullptr, dA.p, dB.p, n->stream);
    h->A = mpcb::Mat(nx, nx); h->B = mpcb::Mat(nx, nu);
    cudaError_t e1 = cudaMemcpyAsync(h->A.a.data(), dA.p, sizeof(double) * nx * nx, cudaMemcpyDeviceToHost, n->stream);
    cudaError_t e2 = cudaMemcpyAsync(h->B.a.data(), dB.p, sizeof(double) * nx * nu, cudaMemcpyDeviceToHost, n->stream);
    cudaError_t e3 = cudaStreamSynchronize(n->stream);
    dx.release(); du.release(); dA.release(); dB.release();
    if (rc != MPCB_OK) { std::string keep = mpcb_last_error(); return bail(rc, keep); }
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) { cudaGetLastError(); return bail(MPCB_ERR_CUDA, "linearisation at the design reference failed"); }
  }
  // host design on the linearisation: P (if absent), rho = sqrt(lmin lmax) of the condensed Hessian
  mpcb_linear_desc ld{};
  ld.nx = nx; ld.nu = nu; ld.horizon = H; ld.A = h->A.a.data(); ld.B = h->B.a.data(); ld.Q = d->Q; ld.R = d->R; ld.S = d->S; ld.P = d->P;
  ld.umin = d->umin; ld.umax = d->umax; ld.xmin = nullptr; ld.xmax = nullptr; ld.state_constraint = 0; ld.terminal_mode = MPCB_TERMINAL_NONE;
  mpcb::Design D;
  std::string err;
  rc = mpcb::build_design(ld, q, D, err);
  if (rc != MPCB_OK) return bail(rc, "design at the reference linearisation: " + err);
  h->rho = D.rho; h->P = D.P;
  // constant Hessian part Hc = 2 (I (x) R) + 2 D'(I (x) S) D with the reference's switches (design_mpc.jl:436-447)
  std::vector<double> Hc((size_t)nz * nz, 0.0), lb(nz), ub(nz);
  if (D.use_R) {
    for (int k = 0; k < H; k++)
      for (int i = 0; i < nu; i++)
        for (int j = 0; j < nu; j++) Hc[(size_t)(k * nu + j) * nz + k * nu + i] += 2.0 * D.R(i, j);
    if (D.use_S)
      for (int k = 0; k + 1 < H; k++)
        for (int i = 0; i < nu; i++)
          for (int j = 0; j < nu; j++) {
            const double s2 = 2.0 * D.S(i, j);
            const int a0 = k * nu + i, a1 = (k + 1) * nu + i, b0 = k * nu + j, b1 = (k + 1) * nu + j;
            Hc[(size_t)b0 * nz + a0] += s2; Hc[(size_t)b1 * nz + a1] += s2; Hc[(size_t)b1 * nz + a0] -= s2; Hc[(size_t)b0 * nz + a1] -= s2;
          }
  }
  if (nu <= 3 * nx) {          // R^-1 for the per-problem Riccati equations of the re-linearised solve (dare.cuh)
    mpcb::Mat Ri = mpcb::Mat::eye(nu);
    if (mpcb::lu_solve(D.R, Ri) && mpcb::upload(h->Rinv, Ri.a.data(), Ri.a.size()) != cudaSuccess) { cudaGetLastError(); return bail(MPCB_ERR_CUDA, "constant upload failed"); }
  }
  for (int e = 0; e < nz; e++) { lb[e] = d->umin[e % nu]; ub[e] = d->umax[e % nu]; if (!(lb[e] <= ub[e])) return bail(MPCB_ERR_INVALID, "umin > umax"); }
  if (mpcb::upload(h->Q, D.Q.a.data(), D.Q.a.size()) != cudaSuccess || mpcb::upload(h->Pt, D.P.a.data(), D.P.a.size()) != cudaSuccess ||
      mpcb::upload(h->Hc, Hc.data(), Hc.size()) != cudaSuccess || mpcb::upload(h->lb, lb.data(), nz) != cudaSuccess ||
      mpcb::upload(h->ub, ub.data(), nz) != cudaSuccess) { cudaGetLastError(); return bail(MPCB_ERR_CUDA, "constant upload failed"); }
  if (h->state_box) {
    for (int i = 0; i < nx; i++) if (!(d->xmin[i] <= d->xmax[i])) return bail(MPCB_ERR_INVALID, "xmin > xmax");
    if (mpcb::upload(h->xmin, d->xmin, nx) != cudaSuccess || mpcb::upload(h->xmax, d->xmax, nx) != cudaSuccess) { cudaGetLastError(); return bail(MPCB_ERR_CUDA, "constant upload failed"); }
  }
  for (auto& e : h->ev)
    if (cudaEventCreate(&e) != cudaSuccess) return bail(MPCB_ERR_CUDA, "cudaEventCreate failed");
  for (size_t i = 1; i < dev_ids.size(); i++) {      // the same controller on the other devices of a multi-device handle
    mpcb_nmpc_settings ps = st;
    ps.qp.n_devices = 0; ps.qp.device = dev_ids[i];
    mpcb_nmpc* peer = nullptr;
    rc = mpcb_create_nmpc(d, &ps, &peer);
    if (rc != MPCB_OK) { const std::string keep = mpcb_last_error(); return bail(rc, "device " + std::to_string(dev_ids[i]) + ": " + keep); }
    h->peers.push_back(peer);
  }
  cudaSetDevice(st.qp.device);
  *out = h;
  return MPCB_OK;
}

void mpcb_destroy_nmpc(mpcb_nmpc* h) {
  if (!h) return;
  for (mpcb_nmpc* p : h->peers) mpcb_destroy_nmpc(p);
  h->peers.clear();
  if (h->nn) { cudaSetDevice(h->nn->device); if (h->nn->stream) cudaStreamSynchronize(h->nn->stream); }
  for (DevBuf<double>* b : {&h->Q, &h->Pt, &h->Hc, &h->lb, &h->ub, &h->xmin, &h->xmax, &h->Rinv, &h->x0, &h->xref, &h->uref, &h->warm_u, &h->warm_y, &h->u, &h->e_u, &h->x, &h->e_x, &h->u0,
                            &h->obj, &h->y, &h->step, &h->dres})
    b->release();
  h->status.release(); h->iters.release(); h->inner.release(); h->counter.release(); h->stage_int.release();
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  mpcb_destroy_nn(h->nn);
  delete h;
}

int mpcb_nmpc_get_design(const mpcb_nmpc* h, double* rho, double* A, double* B, double* P) {
  if (!h) return api_fail(MPCB_ERR_INVALID, "null handle");
  if (rho) *rho = h->rho;
  if (A) std::memcpy(A, h->A.a.data(), sizeof(double) * h->A.a.size());
  if (B) std::memcpy(B, h->B.a.data(), sizeof(double) * h->B.a.size());
  if (P) std::memcpy(P, h->P.a.data(), sizeof(double) * h->P.a.size());
  return MPCB_OK;
}

int mpcb_nmpc_get_timing(const mpcb_nmpc* h, mpcb_timing* t) {
  if (!h || !t) return api_fail(MPCB_ERR_INVALID, "null argument");
  *t = h->timing;
  return MPCB_OK;
}

int mpcb_solve_nmpc_batch_device(mpcb_nmpc* h, const mpcb_batch_io* io, void* cuda_stream) {
  if (!h || !io) return api_fail(MPCB_ERR_INVALID, "null argument");
  if (!h->peers.empty()) return api_fail(MPCB_ERR_INVALID, "multi-device NMPC handles shard through the host entries (mpcb_solve_nmpc_batch); the device entry needs a single-device handle");
  CUDA_TRY(cudaSetDevice(h->nn->device));
  return enqueue_nmpc(h, *io, (cudaStream_t)cuda_stream);
}

int mpcb_solve_relinearized_batch_device(mpcb_nmpc* h, const mpcb_batch_io* io, void* cuda_stream) {
  if (!h || !io) return api_fail(MPCB_ERR_INVALID, "null argument");
  if (!h->peers.empty()) return api_fail(MPCB_ERR_INVALID, "multi-device NMPC handles shard through the host entries (mpcb_solve_relinearized_batch); the device entry needs a single-device handle");
  CUDA_TRY(cudaSetDevice(h->nn->device));
  return enqueue_nmpc(h, *io, (cudaStream_t)cuda_stream, true);
}

static int solve_nmpc_host(mpcb_nmpc* h, const mpcb_batch_io* hio, bool lin);
static int solve_nmpc_host_sharded(mpcb_nmpc* h, const mpcb_batch_io* hio, bool lin) {
  if (!h || !hio) return api_fail(MPCB_ERR_INVALID, "null argument");
  const int ndev = 1 + (int)h->peers.size();
  if (ndev == 1 || hio->batch < mpcb::MULTI_MIN_PER_DEVICE * ndev) return solve_nmpc_host(h, hio, lin);
  if (!hio->x0 || !hio->xref || !hio->uref) return api_fail(MPCB_ERR_INVALID, "x0, xref, uref are required");
  const size_t nx = h->nn->net.nx, nu = h->nn->net.nu, H = h->H, nz = h->nz;
  const mpcb::ShardDims d{nx, nu, H, nz, nz + (h->state_box ? nx * H : 0) + (h->terminal_eq ? nx : 0)};
  int rc = mpcb::run_sharded(ndev, [&](int r) {
    long long lo, hi;
    mpcb::shard_range(hio->batch, r, ndev, &lo, &hi);
    if (hi <= lo) return (int)MPCB_OK;
    const mpcb_batch_io sio = mpcb::shard_io(*hio, lo, hi - lo, d);
    return solve_nmpc_host(r == 0 ? h : h->peers[(size_t)r - 1], &sio, lin);
  });
  if (rc != MPCB_OK) return rc;
  for (mpcb_nmpc* p : h->peers) {
    h->timing.total_ms = std::max(h->timing.total_ms, p->timing.total_ms); h->timing.solve_ms = std::max(h->timing.solve_ms, p->timing.solve_ms);
    h->timing.total_iterations += p->timing.total_iterations; h->timing.kernel_launches += p->timing.kernel_launches;
  }
  h->timing.batch = hio->batch;
  CUDA_TRY(cudaSetDevice(h->nn->device));
  return MPCB_OK;
}
int mpcb_solve_nmpc_batch(mpcb_nmpc* h, const mpcb_batch_io* hio) { return solve_nmpc_host_sharded(h, hio, false); }
int mpcb_solve_relinearized_batch(mpcb_nmpc* h, const mpcb_batch_io* hio) { return solve_nmpc_host_sharded(h, hio, true); }

static int solve_nmpc_host(mpcb_nmpc* h, const mpcb_batch_io* hio, bool lin) {
  if (!h || !hio) return api_fail(MPCB_ERR_INVALID, "null argument");
  const long long Bn = hio->batch;
  if (Bn <= 0) return api_fail(MPCB_ERR_INVALID, "batch must be positive");
  if (!hio->x0 || !hio->xref || !hio->uref) return api_fail(MPCB_ERR_INVALID, "x0, xref, uref are required");
  mpcb_nn* n = h->nn;
  CUDA_TRY(cudaSetDevice(n->device));
  cudaStream_t st = n->stream;
  const size_t nx = n->net.nx, nu = n->net.nu, H = h->H, nz = h->nz, B = (size_t)Bn, ny = nz + (h->state_box ? nx * H : 0) + (h->terminal_eq ? nx : 0);
  const size_t n_xref = hio->xref_broadcast ? nx : nx * B, n_uref = hio->uref_broadcast ? nu : nu * B;
  struct In { const double* src; DevBuf<double>* dst; size_t n; };
  In ins[5] = {{hio->x0, &h->x0, nx * B}, {hio->xref, &h->xref, n_xref}, {hio->uref, &h->uref, n_uref}, {hio->warm_u, &h->warm_u, nz * B}, {hio->warm_y, &h->warm_y, ny * B}};
  CUDA_TRY(cudaEventRecord(h->ev[0], st));
  for (auto& in : ins) {
    if (!in.src) continue;
    CUDA_TRY(in.dst->ensure(in.n));
    CUDA_TRY(cudaMemcpyAsync(in.dst->p, in.src, in.n * sizeof(double), cudaMemcpyHostToDevice, st));   // pageable sources are staged by the driver
  }
  CUDA_TRY(cudaEventRecord(h->ev[1], st));
  mpcb_batch_io dio = *hio;
  dio.x0 = h->x0.p; dio.xref = h->xref.p; dio.uref = h->uref.p;
  dio.warm_u = hio->warm_u ? h->warm_u.p : nullptr; dio.warm_y = hio->warm_y ? h->warm_y.p : nullptr;
  struct Out { double* host; DevBuf<double>* dev; size_t n; double** slot; };
  Out outs[9] = {{hio->u, &h->u, nz * B, &dio.u}, {hio->e_u, &h->e_u, nz * B, &dio.e_u}, {hio->x, &h->x, nx * (H + 1) * B, &dio.x},
                 {hio->e_x, &h->e_x, nx * (H + 1) * B, &dio.e_x}, {hio->u0, &h->u0, nu * B, &dio.u0}, {hio->prim_res, &h->step, B, &dio.prim_res},
                 {hio->dual_res, &h->dres, B, &dio.dual_res}, {hio->objective, &h->obj, B, &dio.objective}, {hio->y, &h->y, ny * B, &dio.y}};
  for (auto& o : outs) {
    if (!o.host) { *o.slot = nullptr; continue; }
    CUDA_TRY(o.dev->ensure(o.n));
    *o.slot = o.dev->p;
  }
  CUDA_TRY(h->status.ensure(B)); CUDA_TRY(h->iters.ensure(B)); CUDA_TRY(h->inner.ensure(B));
  dio.status = h->status.p; dio.iters = h->iters.p; dio.inner_iters = h->inner.p;
  int rc = enqueue_nmpc(h, dio, st, lin);
  if (rc != MPCB_OK) return rc;
  CUDA_TRY(cudaEventRecord(h->ev[2], st));
  for (auto& o : outs)
    if (o.host) CUDA_TRY(cudaMemcpyAsync(o.host, o.dev->p, o.n * sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(h->stage_int.ensure(3 * B));
  CUDA_TRY(cudaMemcpyAsync(h->stage_int.p, h->status.p, B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(h->stage_int.p + B, h->iters.p, B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(h->stage_int.p + 2 * B, h->inner.p, B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaEventRecord(h->ev[3], st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (hio->status) std::memcpy(hio->status, h->stage_int.p, B * sizeof(int32_t));
  if (hio->iters) std::memcpy(hio->iters, h->stage_int.p + B, B * sizeof(int32_t));
  if (hio->inner_iters) std::memcpy(hio->inner_iters, h->stage_int.p + 2 * B, B * sizeof(int32_t));
  long long tot = 0;
  for (size_t i = 0; i < B; i++) tot += h->stage_int.p[2 * B + i];
  h->timing.total_iterations = tot;
  cudaEventElapsedTime(&h->timing.h2d_ms, h->ev[0], h->ev[1]);
  cudaEventElapsedTime(&h->timing.solve_ms, h->ev[1], h->ev[2]);
  h->timing.recover_ms = 0.f;
  cudaEventElapsedTime(&h->timing.d2h_ms, h->ev[2], h->ev[3]);
  cudaEventElapsedTime(&h->timing.total_ms, h->ev[0], h->ev[3]);
  return MPCB_OK;
}

static int closed_loop_nmpc_single(mpcb_nmpc* h, const mpcb_closed_loop_io* cio);
int mpcb_closed_loop_nmpc_batch(mpcb_nmpc* h, const mpcb_closed_loop_io* cio) {
  if (!h || !cio) return api_fail(MPCB_ERR_INVALID, "null argument");
  const int ndev = 1 + (int)h->peers.size();
  if (ndev == 1 || cio->batch < mpcb::MULTI_MIN_PER_DEVICE * ndev) return closed_loop_nmpc_single(h, cio);
  if (!cio->x0 || !cio->xref || !cio->uref) return api_fail(MPCB_ERR_INVALID, "x0, xref, uref are required");
  int rc = mpcb::run_sharded(ndev, [&](int r) {
    long long lo, hi;
    mpcb::shard_range(cio->batch, r, ndev, &lo, &hi);
    if (hi <= lo) return (int)MPCB_OK;
    const mpcb_closed_loop_io sio = mpcb::shard_closed_loop_io(*cio, lo, hi - lo, (size_t)h->nn->net.nx, (size_t)h->nn->net.nu);
    return closed_loop_nmpc_single(r == 0 ? h : h->peers[(size_t)r - 1], &sio);
  });
  if (rc != MPCB_OK) return rc;
  CUDA_TRY(cudaSetDevice(h->nn->device));
  return MPCB_OK;
}

static int closed_loop_nmpc_single(mpcb_nmpc* h, const mpcb_closed_loop_io* cio) {
  const long long Bn = cio->batch;
  const int T = cio->steps;
  if (Bn <= 0 || T <= 0) return api_fail(MPCB_ERR_INVALID, "batch and steps must be positive");
  if (!cio->x0 || !cio->xref || !cio->uref) return api_fail(MPCB_ERR_INVALID, "x0, xref, uref are required");
  mpcb_nn* n = h->nn;
  CUDA_TRY(cudaSetDevice(n->device));
  cudaStream_t st = n->stream;
  const size_t nx = n->net.nx, nu = n->net.nu, H = h->H, nz = h->nz, B = (size_t)Bn, ny = nz + (h->state_box ? nx * H : 0) + (h->terminal_eq ? nx : 0);
  const size_t n_xref = cio->xref_broadcast ? nx : nx * B, n_uref = cio->uref_broadcast ? nu : nu * B;
  const size_t smem = sizeof(double) * (n->net.weight_count() + mpcb::NN_WARPS * mpcb::nn_eval_scratch_doubles(n->net, false));
  if (smem > 200 * 1024) return api_fail(MPCB_ERR_INVALID, "network too large for the shared-memory resident kernels");
  const bool warm = cio->warm_start != 0;
  DevBuf<double> xa, xb, ub, yb, wu, wy, xtraj, utraj;
  DevBuf<int32_t> itot, unsol;
  auto cleanup = [&]() { xa.release(); xb.release(); ub.release(); yb.release(); wu.release(); wy.release(); xtraj.release(); utraj.release(); itot.release(); unsol.release(); };
#define CL_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cudaGetLastError(); cleanup(); return api_fail(MPCB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } } while (0)
  CL_TRY(xa.ensure(nx * B)); CL_TRY(xb.ensure(nx * B)); CL_TRY(ub.ensure(nz * B));
  if (warm) { CL_TRY(yb.ensure(ny * B)); CL_TRY(wu.ensure(nz * B)); CL_TRY(wy.ensure(ny * B)); }
  CL_TRY(h->xref.ensure(n_xref)); CL_TRY(h->uref.ensure(n_uref)); CL_TRY(h->status.ensure(B)); CL_TRY(h->iters.ensure(B)); CL_TRY(h->inner.ensure(B));
  if (cio->x_traj) CL_TRY(xtraj.ensure(nx * (size_t)(T + 1) * B));
  if (cio->u_traj) CL_TRY(utraj.ensure(nu * (size_t)T * B));
  if (cio->iters_total) CL_TRY(itot.ensure(B));
  if (cio->unsolved_steps) CL_TRY(unsol.ensure(B));
  if (smem > 48 * 1024) CL_TRY(cudaFuncSetAttribute(nn_plant_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CL_TRY(cudaMemcpyAsync(xa.p, cio->x0, nx * B * sizeof(double), cudaMemcpyHostToDevice, st));
  CL_TRY(cudaMemcpyAsync(h->xref.p, cio->xref, n_xref * sizeof(double), cudaMemcpyHostToDevice, st));
  CL_TRY(cudaMemcpyAsync(h->uref.p, cio->uref, n_uref * sizeof(double), cudaMemcpyHostToDevice, st));
  CL_TRY(cudaEventRecord(h->ev[0], st));
  int launches = 0;
  double *xc = xa.p, *xn = xb.p;
  const long long blocks = (Bn + mpcb::NN_WARPS - 1) / mpcb::NN_WARPS;
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(blocks, (long long)n->sm_count * 8));
  for (int t = 0; t < T; t++) {
    mpcb_batch_io dio;
    std::memset(&dio, 0, sizeof(dio));
    dio.batch = Bn; dio.x0 = xc; dio.xref = h->xref.p; dio.uref = h->uref.p; dio.xref_broadcast = cio->xref_broadcast; dio.uref_broadcast = cio->uref_broadcast;
    if (warm && t > 0) { dio.warm_u = wu.p; dio.warm_y = wy.p; }
    dio.status = h->status.p; dio.iters = h->iters.p; dio.inner_iters = h->inner.p;
    dio.u = ub.p; dio.y = warm ? yb.p : nullptr;
    int rc = enqueue_nmpc(h, dio, st);
    if (rc != MPCB_OK) { cleanup(); return rc; }
    NnPlantParams Pp{};
    Pp.net = n->net; Pp.H = h->H; Pp.nz = h->nz; Pp.ny = (int)ny; Pp.t = t; Pp.steps = T; Pp.batch = Bn;
    Pp.u = ub.p; Pp.y = warm ? yb.p : nullptr; Pp.status = h->status.p; Pp.inner = h->inner.p; Pp.x_in = xc; Pp.x_out = xn;
    Pp.warm_u = warm ? wu.p : nullptr; Pp.warm_y = warm ? wy.p : nullptr;
    Pp.x_traj = cio->x_traj ? xtraj.p : nullptr; Pp.u_traj = cio->u_traj ? utraj.p : nullptr;
    Pp.iters_total = cio->iters_total ? itot.p : nullptr; Pp.unsolved = cio->unsolved_steps ? unsol.p : nullptr;
    nn_plant_step_kernel<<<grid, mpcb::NN_THREADS, smem, st>>>(Pp);
    CL_TRY(cudaGetLastError());
    launches += 2;
    std::swap(xc, xn);
  }
  CL_TRY(cudaEventRecord(h->ev[2], st));
  if (cio->x_traj) CL_TRY(cudaMemcpyAsync(cio->x_traj, xtraj.p, nx * (size_t)(T + 1) * B * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (cio->u_traj) CL_TRY(cudaMemcpyAsync(cio->u_traj, utraj.p, nu * (size_t)T * B * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (cio->iters_total) CL_TRY(cudaMemcpyAsync(cio->iters_total, itot.p, B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (cio->unsolved_steps) CL_TRY(cudaMemcpyAsync(cio->unsolved_steps, unsol.p, B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CL_TRY(cudaEventRecord(h->ev[3], st));
  CL_TRY(cudaStreamSynchronize(st));
#undef CL_TRY
  cudaEventElapsedTime(&h->timing.solve_ms, h->ev[0], h->ev[2]);
  cudaEventElapsedTime(&h->timing.d2h_ms, h->ev[2], h->ev[3]);
  h->timing.h2d_ms = 0.f; h->timing.recover_ms = 0.f; h->timing.total_ms = h->timing.solve_ms + h->timing.d2h_ms;
  h->timing.batch = Bn; h->timing.kernel_launches = launches; h->timing.chunks = 1; h->timing.total_iterations = 0;
  cleanup();
  return MPCB_OK;
}

}  // extern "C"
