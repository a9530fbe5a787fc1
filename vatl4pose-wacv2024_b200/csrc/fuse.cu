// Score fusion of active_learning/ActiveLearning.py:490-516 in float64 on the device:
// per-criterion min-max over the unlabelled rows, combination, min-max again.  Every
// operation is a single IEEE double op in the reference's order, so given the same inputs
// the result is bit-identical to numpy's (NaN inputs included: they poison the statistics like np.min / np.max).
#include "common.cuh"

namespace vatlq {

// IEEE min via compare-and-swap (min is order independent -> deterministic)
__device__ __forceinline__ void atomic_min_f64(double* addr, double val) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a;
  while (true) {
    const double cur = __longlong_as_double((long long)old);
    if (cur <= val) return;
    const unsigned long long prev = atomicCAS(a, old, (unsigned long long)__double_as_longlong(val));
    if (prev == old) return;
    old = prev;
  }
}

__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <int NV>
__device__ __forceinline__ void block_min_to(double (&v)[NV], double* out) {
  __shared__ double s[NV][8];
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = warp_min(v[k]);
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int k = 0; k < NV; ++k) s[k][threadIdx.x >> 5] = v[k];
  __syncthreads();
  if (threadIdx.x < NV) {
    double m = s[threadIdx.x][0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmin(m, s[threadIdx.x][w]);
    atomic_min_f64(out + threadIdx.x, m);
  }
}

__global__ void __launch_bounds__(256) fuse_stats_kernel(const float* __restrict__ thc, const float* __restrict__ wpu,
                                                         const uint8_t* __restrict__ unl, long long n, double* stats) {
  double v[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (unl && !unl[i]) continue;
    // np.min / np.max propagate NaN (fmin drops it): a NaN input sets (min, -max) = (-inf, -inf), which turns
    // every normalised score into NaN exactly like the reference's (x - min) / (max - min)
    const double t = (double)thc[i];
    if (t != t) v[0] = v[1] = -INFINITY;
    v[0] = fmin(v[0], t);
    v[1] = fmin(v[1], -t);
    if (wpu) {
      const double w = (double)wpu[i];
      if (w != w) v[2] = v[3] = -INFINITY;
      v[2] = fmin(v[2], w);
      v[3] = fmin(v[3], -w);
    }
  }
  block_min_to<4>(v, stats);
}

__device__ __forceinline__ double minmax1(double x, double mn, double mx) {
  return __ddiv_rn(__dsub_rn(x, mn), __dsub_rn(mx, mn));
}

__global__ void __launch_bounds__(256) fuse_combine_kernel(const float* __restrict__ thc, const float* __restrict__ wpu,
                                                           const uint8_t* __restrict__ unl, long long n,
                                                           const double* __restrict__ st, int mode, double ratio,
                                                           double* __restrict__ u, double* stats2) {
  const double tmin = st[0], tmax = -st[1], wmin = st[2], wmax = -st[3];
  double v[2] = {INFINITY, INFINITY};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (unl && !unl[i]) {
      u[i] = 0.0;
      continue;
    }
    double r;
    if (mode == 3) {
      r = (double)thc[i];  // single criterion: one min-max, applied by fuse_final (:511-516)
    } else {
      const double a = minmax1((double)thc[i], tmin, tmax);  // :497
      const double b = minmax1((double)wpu[i], wmin, wmax);  // :498
      if (mode == 0) r = __dadd_rn(a, b);                                                        // :501
      else if (mode == 1) r = __dadd_rn(__dmul_rn(ratio, a), __dmul_rn(__dsub_rn(1.0, ratio), b));  // :505
      else r = __dadd_rn(__dmul_rn(__dsub_rn(1.0, ratio), a), __dmul_rn(ratio, b));                // :507
    }
    u[i] = r;
    if (r != r) v[0] = v[1] = -INFINITY;
    v[0] = fmin(v[0], r);
    v[1] = fmin(v[1], -r);
  }
  block_min_to<2>(v, stats2);
}

__global__ void __launch_bounds__(256) fuse_final_kernel(const uint8_t* __restrict__ unl, long long n,
                                                         const double* __restrict__ st2, double* __restrict__ u) {
  const double mn = st2[0], mx = -st2[1];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (unl && !unl[i]) u[i] = 0.0;
    else u[i] = minmax1(u[i], mn, mx);  // :509 / :515
  }
}

static unsigned grid_for(long long n) {
  long long g = (n + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace vatlq

using namespace vatlq;

extern "C" int vatlq_fuse_stats(const float* thc, const float* wpu, const uint8_t* unlabeled, int64_t n,
                                double* stats1, vatlq_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  VQ_REQUIRE(thc && stats1 && n >= 0, "null pointer");
  if (int e = fill_f64(stats1, 4, INFINITY, stream)) return e;
  if (n == 0) return 0;
  fuse_stats_kernel<<<grid_for(n), 256, 0, stream>>>(thc, wpu, unlabeled, n, stats1);
  VQ_LAUNCHED();
  return 0;
}

extern "C" int vatlq_fuse_combine(const float* thc, const float* wpu, const uint8_t* unlabeled, int64_t n,
                                  const double* stats1, int mode, double labeled_ratio, double* u,
                                  double* stats2, vatlq_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  VQ_REQUIRE(thc && stats1 && u && stats2 && n >= 0, "null pointer");
  VQ_REQUIRE(mode >= 0 && mode <= 3, "mode must be 0..3");
  VQ_REQUIRE(mode == 3 || wpu != nullptr, "two-criterion modes need wpu");
  if (int e = fill_f64(stats2, 2, INFINITY, stream)) return e;
  if (n == 0) return 0;
  fuse_combine_kernel<<<grid_for(n), 256, 0, stream>>>(thc, wpu, unlabeled, n, stats1, mode, labeled_ratio, u, stats2);
  VQ_LAUNCHED();
  return 0;
}

extern "C" int vatlq_fuse_final(const uint8_t* unlabeled, int64_t n, const double* stats2, double* u_inout,
                                vatlq_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  VQ_REQUIRE(stats2 && u_inout && n >= 0, "null pointer");
  if (n == 0) return 0;
  fuse_final_kernel<<<grid_for(n), 256, 0, stream>>>(unlabeled, n, stats2, u_inout);
  VQ_LAUNCHED();
  return 0;
}
