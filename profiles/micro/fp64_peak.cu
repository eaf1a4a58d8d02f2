// FP64 peak micro-benchmarks for B200 (sm_100a): MEASURED_PEAKS.json carries no FP64 number,
// so the roofline denominator for the ADMM kernels is measured here.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo fp64_peak.cu -o fp64_peak.bin
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CH>
__global__ void k_dfma(double* out, int iters, double s) {
  double acc[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) acc[i] = threadIdx.x * 1e-3 + i;
  double m = 1.0 + s * 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) acc[i] = fma(acc[i], m, s);
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) r += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int CH>
__global__ void k_dmma(double* out, int iters, double s) {
  double c0[CH], c1[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) { c0[i] = threadIdx.x * 1e-3; c1[i] = i; }
  double a = 1e-3 * (threadIdx.x & 3) + s, b = 1e-3 * (threadIdx.x >> 2) + s;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) dmma(c0[i], c1[i], a, b);
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) r += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// ADMM-shaped loop: NT=40 -> 10 k-steps x 5 n-tiles, B fragment streamed from shared memory,
// A fragment produced from registers with a few DFMAs (EW = elementwise fp64 ops per k-step).
template <int KS, int NTL, int EW>
__global__ void k_dmma_lds(double* out, const double* Tg, int iters, double s) {
  extern __shared__ double Ts[];
  for (int i = threadIdx.x; i < KS * NTL * 32; i += blockDim.x) Ts[i] = Tg[i];
  __syncthreads();
  int lane = threadIdx.x & 31;
  double st[KS];
#pragma unroll
  for (int k = 0; k < KS; k++) st[k] = 1e-3 * (lane + k);
  double r = 0;
  for (int it = 0; it < iters; it++) {
    double c0[NTL], c1[NTL];
#pragma unroll
    for (int t = 0; t < NTL; t++) { c0[t] = 0; c1[t] = 0; }
#pragma unroll
    for (int k = 0; k < KS; k++) {
      double a = st[k];
#pragma unroll
      for (int e = 0; e < EW; e++) a = fma(a, 0.999, s);
#pragma unroll
      for (int t = 0; t < NTL; t++) dmma(c0[t], c1[t], a, Ts[(k * NTL + t) * 32 + lane]);
    }
#pragma unroll
    for (int t = 0; t < NTL; t++) { st[2 * t] = c0[t] * 1e-3; st[2 * t + 1] = c1[t] * 1e-3; }
  }
#pragma unroll
  for (int k = 0; k < KS; k++) r += st[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <typename F>
float timeit(F f, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 64 * 1024));
  double* Tg; CK(cudaMalloc(&Tg, sizeof(double) * 64 * 64)); CK(cudaMemset(Tg, 0, sizeof(double) * 64 * 64));
  const int iters = 20000;
  for (int thr : {128, 256, 512, 1024}) for (int bps : {1, 2}) {
    if (thr * bps > 2048) continue;
    int grid = sms * bps;
    float ms = timeit([&] { k_dfma<8><<<grid, thr>>>(out, iters, 1e-9); });
    double fl = 2.0 * 8 * iters * (double)grid * thr;
    printf("{\"bench\": \"dfma\", \"threads\": %d, \"blocks_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", thr, bps, ms, fl / ms / 1e9);
  }
  for (int thr : {128, 256, 512, 1024}) for (int bps : {1, 2}) {
    if (thr * bps > 2048) continue;
    int grid = sms * bps;
    float ms = timeit([&] { k_dmma<8><<<grid, thr>>>(out, iters / 4, 1e-9); });
    double fl = 2.0 * 256 * 8 * (iters / 4) * (double)grid * (thr / 32);
    printf("{\"bench\": \"dmma_regs_ch8\", \"threads\": %d, \"blocks_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", thr, bps, ms, fl / ms / 1e9);
    ms = timeit([&] { k_dmma<2><<<grid, thr>>>(out, iters / 4, 1e-9); });
    fl = 2.0 * 256 * 2 * (iters / 4) * (double)grid * (thr / 32);
    printf("{\"bench\": \"dmma_regs_ch2\", \"threads\": %d, \"blocks_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", thr, bps, ms, fl / ms / 1e9);
  }
  // single-warp dependent chain latency
  {
    float ms = timeit([&] { k_dmma<1><<<1, 32>>>(out, 100000, 1e-9); });
    printf("{\"bench\": \"dmma_latency_1warp_chain\", \"ns_per_dmma\": %.3f}\n", ms * 1e6 / 100000);
    ms = timeit([&] { k_dfma<1><<<1, 32>>>(out, 100000, 1e-9); });
    printf("{\"bench\": \"dfma_latency_1warp_chain\", \"ns_per_dfma\": %.3f}\n", ms * 1e6 / 100000);
  }
  for (int thr : {128, 256, 384, 512}) for (int bps : {1, 2, 4}) {
    if (thr * bps > 2048) continue;
    int grid = sms * bps;
    size_t sh = 10 * 5 * 32 * sizeof(double);
    float ms = timeit([&] { k_dmma_lds<10, 5, 0><<<grid, thr, sh>>>(out, Tg, 2000, 1e-9); });
    double fl = 2.0 * 256 * 50 * 2000 * (double)grid * (thr / 32);
    printf("{\"bench\": \"dmma_lds_nt40_ew0\", \"threads\": %d, \"blocks_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", thr, bps, ms, fl / ms / 1e9);
    ms = timeit([&] { k_dmma_lds<10, 5, 12><<<grid, thr, sh>>>(out, Tg, 2000, 1e-9); });
    printf("{\"bench\": \"dmma_lds_nt40_ew12\", \"threads\": %d, \"blocks_per_sm\": %d, \"ms\": %.4f, \"tflops_mma_only\": %.3f}\n", thr, bps, ms, fl / ms / 1e9);
  }
  return 0;
}
