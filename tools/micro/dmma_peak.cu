// Microbenchmark: issue rate of mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) from CUDA C on sm_100a as a
// function of the register-tile shape (MI x NJ accumulators per warp, distinct A per row and B
// per column, i-major issue order) and of the warps per SM.  The denominator and the design
// guide for the dense-generator path (csrc/dense_mma.cuh; tools/bench_configs.py --configs 5).
#include <cstdio>
#include <cuda_runtime.h>

template <int MI, int NJ>
__global__ void __launch_bounds__(512) k(double* out, const double* in, int iters) {
  double acc[MI][NJ][2];
  double a[MI], b[NJ];
#pragma unroll
  for (int i = 0; i < MI; ++i) a[i] = in[threadIdx.x + 32 * i];
#pragma unroll
  for (int j = 0; j < NJ; ++j) b[j] = in[threadIdx.x + 32 * (MI + j)];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                     : "+d"(acc[i][j][0]), "+d"(acc[i][j][1]) : "d"(a[i]), "d"(b[j]));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) s += acc[i][j][0] + acc[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MI, int NJ>
void run(int warps_per_sm, double* d_out, const double* d_in) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 200000 / (MI * NJ);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MI, NJ><<<sms, warps_per_sm * 32>>>(d_out, d_in, 100);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MI, NJ><<<sms, warps_per_sm * 32>>>(d_out, d_in, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double flops = (double)sms * warps_per_sm * iters * MI * NJ * 512.0;
  printf("{\"tile\": \"%dx%d\", \"warps_per_sm\": %d, \"tflops\": %.2f}\n", MI, NJ, warps_per_sm, flops / ms / 1e9);
}

int main() {
  double *d_out, *d_in;
  cudaMalloc(&d_out, sizeof(double) * 148 * 1024 * 2);
  cudaMalloc(&d_in, sizeof(double) * 4096);
  cudaMemset(d_in, 0, sizeof(double) * 4096);
  for (int w : {8, 16}) {  // <= 512 threads: 128 registers per thread, no spills up to 14x2
    run<1, 1>(w, d_out, d_in);
    run<1, 2>(w, d_out, d_in);
    run<2, 2>(w, d_out, d_in);
    run<4, 2>(w, d_out, d_in);
    run<7, 2>(w, d_out, d_in);
    run<2, 4>(w, d_out, d_in);
    run<4, 4>(w, d_out, d_in);
    run<1, 8>(w, d_out, d_in);
    run<8, 1>(w, d_out, d_in);
    run<2, 8>(w, d_out, d_in);
    run<7, 1>(w, d_out, d_in);
    run<14, 2>(w, d_out, d_in);
  }
  return 0;
}
