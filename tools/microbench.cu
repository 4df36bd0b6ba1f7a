// microbench.cu -- FP64 pipe measurements on B200 that decide the K1 formulation (DESIGN.md "K1"):
//   dfma   : vector DFMA throughput vs warps/SM and independent chains per thread
//   dmma   : mma.sync.m8n8k4.f64 (DMMA) throughput vs warps/SM and independent accumulators
//   mix    : DMMA and DFMA interleaved in one warp -- do they share one pipe?
//   lds    : shared-memory broadcast-read cost next to DFMA (the operand-delivery bound of the DFMA form)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/microbench tools/microbench.cu
// Output: one JSON line per measurement on stdout.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x)                                                                  \
  do {                                                                         \
    cudaError_t e_ = (x);                                                      \
    if (e_ != cudaSuccess) {                                                   \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                 \
      exit(1);                                                                 \
    }                                                                          \
  } while (0)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

template <int ILP> __global__ void k_dfma(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP> __global__ void k_dmma(double* out, int iters, double a, double b) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) c0[i] = threadIdx.x + i, c1[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int i = 0; i < ILP; i++) dmma(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c0[i] + c1[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// per inner step: ND DMMAs (independent accumulators) + NF DFMAs (independent chains)
template <int ND, int NF> __global__ void k_mix(double* out, int iters, double a, double b) {
  double c0[ND > 0 ? ND : 1], c1[ND > 0 ? ND : 1], x[NF > 0 ? NF : 1];
#pragma unroll
  for (int i = 0; i < ND; i++) c0[i] = threadIdx.x + i, c1[i] = i;
#pragma unroll
  for (int i = 0; i < NF; i++) x[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < ND; i++) dmma(c0[i], c1[i], a, b);
#pragma unroll
      for (int i = 0; i < NF; i++) x[i] = fma(x[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ND; i++) s += c0[i] + c1[i];
#pragma unroll
  for (int i = 0; i < NF; i++) s += x[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// NL shared-memory loads of width W bytes per NF DFMAs; quarter-warp-broadcast addressing (8 lanes share an address)
template <int NL, int NF, int W> __global__ void k_lds(double* out, int iters, double a) {
  __shared__ __align__(16) double sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  double x[NF];
#pragma unroll
  for (int i = 0; i < NF; i++) x[i] = threadIdx.x + i;
  const int grp = (threadIdx.x & 31) >> 3;
  int off = grp * 66 + (threadIdx.x >> 5) * 264;
  for (int it = 0; it < iters; it++) {
    double v[NL * (W / 8)];
#pragma unroll
    for (int l = 0; l < NL; l++) {
      if (W == 16) {
        const double2 t = *reinterpret_cast<const double2*>(&sm[(off + 2 * l) & 4094]);
        v[2 * l] = t.x, v[2 * l + 1] = t.y;
      } else {
        v[l] = sm[(off + l) & 4095];
      }
    }
#pragma unroll
    for (int i = 0; i < NF; i++) x[i] = fma(x[i], a, v[i % (NL * (W / 8))]);
    off += 2;
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NF; i++) s += x[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static double* g_out;
static int g_sms;

template <class F> static float time_ms(F launch) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep >= 1 && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return best;
}

template <int ILP> static void run_dfma(int warps_per_sm) {
  const int threads = 128, blocks = g_sms * warps_per_sm / 4, iters = 2048;
  float ms = time_ms([&] { k_dfma<ILP><<<blocks, threads>>>(g_out, iters, 0.999999, 1e-9); });
  const double fma_n = (double)blocks * threads * iters * 8.0 * ILP;
  printf("{\"test\":\"dfma\",\"warps_per_sm\":%d,\"ilp\":%d,\"ms\":%.4f,\"tflops\":%.3f,\"fma_per_clk_sm\":%.2f}\n", warps_per_sm,
         ILP, ms, 2 * fma_n / ms / 1e9, fma_n / (ms * 1e-3) / g_sms / 1.965e9);
}
template <int ILP> static void run_dmma(int warps_per_sm) {
  const int threads = 128, blocks = g_sms * warps_per_sm / 4, iters = 512;
  float ms = time_ms([&] { k_dmma<ILP><<<blocks, threads>>>(g_out, iters, 0.999999, 1e-9); });
  const double fma_n = (double)blocks * (threads / 32) * iters * 8.0 * ILP * 256.0;
  printf("{\"test\":\"dmma\",\"warps_per_sm\":%d,\"ilp\":%d,\"ms\":%.4f,\"tflops\":%.3f,\"fma_per_clk_sm\":%.2f}\n", warps_per_sm,
         ILP, ms, 2 * fma_n / ms / 1e9, fma_n / (ms * 1e-3) / g_sms / 1.965e9);
}
template <int ND, int NF> static void run_mix(int warps_per_sm) {
  const int threads = 128, blocks = g_sms * warps_per_sm / 4, iters = 512;
  float ms = time_ms([&] { k_mix<ND, NF><<<blocks, threads>>>(g_out, iters, 0.999999, 1e-9); });
  const double n_dmma = (double)blocks * (threads / 32) * iters * 4.0 * ND * 256.0;
  const double n_dfma = (double)blocks * threads * iters * 4.0 * NF;
  printf("{\"test\":\"mix\",\"warps_per_sm\":%d,\"dmma\":%d,\"dfma\":%d,\"ms\":%.4f,\"dmma_fma_per_clk_sm\":%.2f,\"dfma_fma_per_clk_sm\":%.2f}\n",
         warps_per_sm, ND, NF, ms, n_dmma / (ms * 1e-3) / g_sms / 1.965e9, n_dfma / (ms * 1e-3) / g_sms / 1.965e9);
}
template <int NL, int NF, int W> static void run_lds(int warps_per_sm) {
  const int threads = 128, blocks = g_sms * warps_per_sm / 4, iters = 8192;
  float ms = time_ms([&] { k_lds<NL, NF, W><<<blocks, threads>>>(g_out, iters, 0.999999); });
  const double n_lds = (double)blocks * (threads / 32) * iters * NL;
  const double n_dfma = (double)blocks * threads * iters * NF;
  printf("{\"test\":\"lds\",\"warps_per_sm\":%d,\"width\":%d,\"lds_per_step\":%d,\"dfma_per_step\":%d,\"ms\":%.4f,\"lds_warpinst_per_clk_sm\":%.3f,\"dfma_fma_per_clk_sm\":%.2f}\n",
         warps_per_sm, W, NL, NF, ms, n_lds / (ms * 1e-3) / g_sms / 1.965e9, n_dfma / (ms * 1e-3) / g_sms / 1.965e9);
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  g_sms = p.multiProcessorCount;
  CK(cudaMalloc(&g_out, sizeof(double) * g_sms * 64 * 128));
  printf("{\"device\":\"%s\",\"sms\":%d,\"clock_khz\":%d}\n", p.name, g_sms, p.clockRate);
  for (int w : {4, 8, 16, 32}) {
    run_dfma<1>(w);
    run_dfma<2>(w);
    run_dfma<4>(w);
    run_dfma<8>(w);
  }
  for (int w : {4, 8, 16, 32}) {
    run_dmma<1>(w);
    run_dmma<2>(w);
    run_dmma<4>(w);
    run_dmma<8>(w);
  }
  for (int w : {8, 16, 32}) {
    run_mix<4, 0>(w);
    run_mix<0, 32>(w);
    run_mix<4, 32>(w);  // equal FMA counts per lane: 4 DMMA = 32 lane-FMAs
    run_mix<4, 16>(w);
    run_mix<2, 32>(w);
  }
  for (int w : {8, 16}) {
    run_lds<4, 16, 16>(w);
    run_lds<4, 8, 16>(w);
    run_lds<8, 16, 8>(w);
    run_lds<8, 8, 8>(w);
    run_lds<1, 16, 16>(w);
  }
  return 0;
}
