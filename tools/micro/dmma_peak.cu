// Microbenchmark: peak issue rate of mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) from CUDA C on sm_100a,
// as a function of independent accumulator chains per warp and warps per SM.  The denominator
// for the dense-generator path (tools/bench_configs.py --configs 5).
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void __launch_bounds__(1024) k(double* out, int iters, double a0, double b0) {
  double acc[NACC][2];
  for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = 0.0;
  double a = a0 + threadIdx.x, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                   : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
void run(int warps_per_sm, double* d_out) {
  int dev_sms = 148;
  cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<NACC><<<dev_sms, warps_per_sm * 32>>>(d_out, 100, 1.0, 2.0);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<NACC><<<dev_sms, warps_per_sm * 32>>>(d_out, iters, 1.0, 2.0);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double flops = (double)dev_sms * warps_per_sm * iters * NACC * 512.0;
  printf("{\"nacc\": %d, \"warps_per_sm\": %d, \"tflops\": %.2f}\n", NACC, warps_per_sm, flops / ms / 1e9);
}

int main() {
  double* d_out;
  cudaMalloc(&d_out, sizeof(double) * 148 * 1024 * 2);
  for (int w : {4, 8, 16, 32}) {
    run<1>(w, d_out);
    run<2>(w, d_out);
    run<4>(w, d_out);
    run<8>(w, d_out);
    run<16>(w, d_out);
    run<28>(w, d_out);
  }
  return 0;
}
