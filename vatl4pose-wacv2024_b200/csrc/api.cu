// Library-wide state and small utilities of libvatlq.
#include "common.cuh"

namespace vatlq {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

__global__ void fill_f64_kernel(double* p, long long n, double v) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
int fill_f64(double* p, long long n, double v, cudaStream_t stream) {
  if (n <= 0) return 0;
  fill_f64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p, n, v);
  VQ_LAUNCHED();
  return 0;
}
}  // namespace vatlq

extern "C" int vatlq_abi_version(void) { return VATLQ_ABI_VERSION; }
extern "C" const char* vatlq_last_error(void) { return vatlq::g_err; }
extern "C" uint64_t vatlq_launch_count(void) { return vatlq::g_launches.load(); }
