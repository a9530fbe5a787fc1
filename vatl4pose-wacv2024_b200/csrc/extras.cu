// Round-2 rows next to the hot path:
//   * OKS of every item against its ground-truth pose (active_learning/al_metric.py:42-69): the
//     controller derives moks_queried (the core-set's score weights, ActiveLearning.py:815-821,858)
//     and the stopping criteria (:707-725) from it.
#include "common.cuh"

namespace vatlq {

// np.add.reduce over n contiguous fp64 values, n < 128 (numpy's pairwise sum: sequential below 8,
// else eight strided accumulators over the leading multiple of 8, fixed combine tree, then the tail)
__device__ __forceinline__ double np_sum_f64(const double* a, int n) {
  if (n < 8) {
    double res = 0.0;   // (numpy starts from a[0]; 0 + a[0] is exact)
    for (int i = 0; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
  }
  double r[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) r[k] = a[k];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], a[i + k]);
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __dadd_rn(res, a[i]);
  return res;
}

__constant__ double c_oks_vars[17];

// one thread per item; fp64 like the reference (python floats / numpy float64)
__global__ void __launch_bounds__(128) oks_kernel(const float* __restrict__ kpts, const float* __restrict__ gt,
                                                  const float* __restrict__ bbox_xyxy, long long n, double* __restrict__ oks) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* d = kpts + i * 51;
  const float* g = gt + i * 51;
  const float* b = bbox_xyxy + i * 4;
  // bbox_xyxy_to_xywh (alphapose/utils/bbox.py:91-97)
  const double bx = b[0], by = b[1];
  const double bw = __dadd_rn(__dsub_rn((double)b[2], (double)b[0]), 1.0);
  const double bh = __dadd_rn(__dsub_rn((double)b[3], (double)b[1]), 1.0);
  int k1 = 0;
  for (int j = 0; j < 17; ++j) k1 += (g[3 * j + 2] > 0.f) ? 1 : 0;
  const double x0 = __dsub_rn(bx, bw), x1 = __dadd_rn(bx, __dmul_rn(bw, 2.0));
  const double y0 = __dsub_rn(by, bh), y1 = __dadd_rn(by, __dmul_rn(bh, 2.0));
  const double area = __dadd_rn(__dmul_rn(bw, bh), 2.220446049250313e-16);   // + np.spacing(1)
  double e[17];
  int m = 0;
  for (int j = 0; j < 17; ++j) {
    const double xd = d[3 * j], yd = d[3 * j + 1];
    double dx, dy;
    if (k1 > 0) {
      dx = __dsub_rn(xd, (double)g[3 * j]);
      dy = __dsub_rn(yd, (double)g[3 * j + 1]);
    } else {
      dx = __dadd_rn(fmax(0.0, __dsub_rn(x0, xd)), fmax(0.0, __dsub_rn(xd, x1)));
      dy = __dadd_rn(fmax(0.0, __dsub_rn(y0, yd)), fmax(0.0, __dsub_rn(yd, y1)));
    }
    double v = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    v = __ddiv_rn(v, c_oks_vars[j]);
    v = __ddiv_rn(v, area);
    v = __dmul_rn(v, 0.5);
    if (k1 == 0 || g[3 * j + 2] > 0.f) e[m++] = exp(-v);
  }
  oks[i] = __ddiv_rn(np_sum_f64(e, m), (double)m);
}

}  // namespace vatlq

using namespace vatlq;

extern "C" int vatlq_oks(const float* kpts, const float* gt_kpts, const float* bbox_ann_xyxy, int64_t n, double* oks,
                         vatlq_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  VQ_REQUIRE(n >= 0, "n must be >= 0");
  if (n == 0) return 0;
  VQ_REQUIRE(kpts && gt_kpts && bbox_ann_xyxy && oks, "null pointer");
  static bool init = false;
  if (!init) {   // OKS_vars = (OKS_sigmas * 2) ** 2, OKS_sigmas = [...] / 10.0   (al_metric.py:38-39)
    const double s[17] = {.26, .25, .25, .35, .35, .79, .79, .72, .72, .62, .62, 1.07, 1.07, .87, .87, .89, .89};
    double v[17];
    for (int j = 0; j < 17; ++j) {
      const double t = (s[j] / 10.0) * 2.0;
      v[j] = t * t;
    }
    VQ_CUDA(cudaMemcpyToSymbol(c_oks_vars, v, sizeof(v)));
    init = true;
  }
  oks_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(kpts, gt_kpts, bbox_ann_xyxy, (long long)n, oks);
  VQ_LAUNCHED();
  return 0;
}
