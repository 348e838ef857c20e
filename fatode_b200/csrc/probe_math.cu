// Test probe: the correctly rounded elementary functions of portable_math.h evaluated ON THE DEVICE, so that tests can
// compare them bit for bit with the host evaluation of the same header (oracle side) and with a multi-precision reference.
#include <cuda_runtime.h>
#include <string>
#include "../../include/fatode_b200.h"
#include "portable_math.h"

namespace {
__global__ void k_probe_math(int n, const double* __restrict__ x, double elo, double* __restrict__ out_cos, double* __restrict__ out_root) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out_cos[i] = fpm::cos_cr(x[i]);
  out_root[i] = fpm::root_elo(fabs(x[i]), elo);
}
}  // namespace

// x[n] host array; out_cos[i] = cos_cr(x[i]) (|x| <= 5 pi/4), out_root[i] = |x[i]|**fl(1/elo)
extern "C" int fatode_probe_math(int n, const double* x, double elo, double* out_cos, double* out_root) {
  if (n <= 0 || !x || !out_cos || !out_root) return FATODE_EINVAL;
  double *d_x = nullptr, *d_c = nullptr, *d_r = nullptr;
  const size_t nb = sizeof(double) * (size_t)n;
  int rc = 0;
  if (cudaMalloc(&d_x, nb) != cudaSuccess || cudaMalloc(&d_c, nb) != cudaSuccess || cudaMalloc(&d_r, nb) != cudaSuccess) rc = FATODE_ENOMEM;
  if (!rc && cudaMemcpy(d_x, x, nb, cudaMemcpyHostToDevice) != cudaSuccess) rc = FATODE_ECUDA;
  if (!rc) {
    k_probe_math<<<(n + 255) / 256, 256>>>(n, d_x, elo, d_c, d_r);
    if (cudaGetLastError() != cudaSuccess || cudaMemcpy(out_cos, d_c, nb, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(out_root, d_r, nb, cudaMemcpyDeviceToHost) != cudaSuccess)
      rc = FATODE_ECUDA;
  }
  cudaFree(d_x); cudaFree(d_c); cudaFree(d_r);
  return rc;
}
