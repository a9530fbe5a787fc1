// Micro-benchmark: peak fp64 FMA rate on this GPU, scalar DFMA vs tensor-core DMMA (mma.sync.m8n8k4.f64)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rate fp64_rate.cu && ./fp64_rate
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = threadIdx.x * 1e-9 + k;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = fma(acc[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void dmma_kernel(double* out, int iters, double a, double b) {
  double c[8][2];
#pragma unroll
  for (int k = 0; k < 8; ++k) { c[k][0] = threadIdx.x * 1e-9; c[k][1] = k; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) dmma(c[k][0], c[k][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void cvt_kernel(double* out, const float* in, int iters) {
  float f[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) f[k] = in[threadIdx.x + k];
  double s = 0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) { s += (double)f[k]; f[k] = __int_as_float(__float_as_int(f[k]) + 1); }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// DMMA with one fp32->fp64 conversion per DMMA (the ratio of the batched core-set pass at 8 centers)
template <int MODE>  // 0: F2F cvt, 1: integer bit conversion
__global__ void dmma_cvt_kernel(double* out, const float* in, int iters, double b) {
  double c[8][2];
  float f[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { c[k][0] = threadIdx.x * 1e-9; c[k][1] = k; f[k] = in[threadIdx.x + k] + 1.0f + k; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      double a;
      if (MODE == 0) a = (double)f[k];
      else {
        const unsigned u = __float_as_uint(f[k]);
        unsigned hi = (((u & 0x7fffffffu) >> 3) + 0x38000000u) | (u & 0x80000000u);
        if ((u & 0x7f800000u) == 0) hi = u & 0x80000000u;
        a = __hiloint2double((int)hi, (int)(u << 29));
      }
      dmma(c[k][0], c[k][1], a, b);
      f[k] = __uint_as_float(__float_as_uint(f[k]) + 3);
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; float* in;
  cudaMalloc(&out, sizeof(double) * sms * 1024 * 4);
  cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0, 4096 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int threads : {256, 512, 1024}) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      dfma_kernel<<<sms * 2, threads>>>(out, iters, 1.0000001, 1e-9);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma = (double)sms * 2 * threads * 8.0 * iters;
      if (rep) printf("DFMA threads=%4d  %.2f T fma/s (%.1f TFLOP/s)  %.3f ms\n", threads, fma / ms / 1e9, 2 * fma / ms / 1e9, ms);
      cudaEventRecord(e0);
      dmma_kernel<<<sms * 2, threads>>>(out, iters, 1.0000001, 1e-9);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      fma = (double)sms * 2 * (threads / 32) * 8.0 * iters * 256.0;
      if (rep) printf("DMMA threads=%4d  %.2f T fma/s (%.1f TFLOP/s)  %.3f ms\n", threads, fma / ms / 1e9, 2 * fma / ms / 1e9, ms);
      cudaEventRecord(e0);
      cvt_kernel<<<sms * 2, threads>>>(out, in, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      double cv = (double)sms * 2 * threads * 8.0 * iters;
      if (rep) printf("CVT+DADD threads=%4d  %.2f T cvt/s  %.3f ms\n", threads, cv / ms / 1e9, ms);
      for (int mode = 0; mode < 2; ++mode) {
        cudaEventRecord(e0);
        if (mode == 0) dmma_cvt_kernel<0><<<sms * 2, threads>>>(out, in, iters, 1e-9);
        else dmma_cvt_kernel<1><<<sms * 2, threads>>>(out, in, iters, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        fma = (double)sms * 2 * (threads / 32) * 8.0 * iters * 256.0;
        if (rep) printf("DMMA+%s 1:1 threads=%4d  %.2f T fma/s (%.1f TFLOP/s)  %.3f ms\n", mode ? "intcvt" : "F2F", threads, fma / ms / 1e9, 2 * fma / ms / 1e9, ms);
      }
    }
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
