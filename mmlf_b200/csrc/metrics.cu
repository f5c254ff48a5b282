// Distribution metrics of the validation script on the GPU: the reference pulls the
// (70 | 108, 512, 512) posteriors to the host and evaluates them with numpy (validate/cli.py:74-118, 174-187, 51-70);
// the ESE case alone is 70 x 109 Laplace-CDF evaluations per pixel in float64.  Same float64 arithmetic here, one thread
// per pixel, the ensemble members of a block staged in shared memory.
#include <math.h>

#include "../../include/mmlf_b200.h"
#include "common.cuh"
#include "host_util.h"

namespace mmlf {

constexpr int kMetThreads = 64;

// out[b][c][px] = 1/K * sum_k (cdf_k(edge_{c+1}) - cdf_k(edge_c)), edges = linspace(x_min - step/2, x_max + step/2, n_bins + 1)
__global__ void __launch_bounds__(kMetThreads)
lmm_to_discrete_kernel(const float* __restrict__ means, const float* __restrict__ logvars, int K, int64_t B, int64_t HW,
                       int n_bins, double first_edge, double edge_step, double last_edge, double* __restrict__ out) {
  extern __shared__ double met_smem[];               // [K][threads] mean, then [K][threads] var
  double* s_mean = met_smem;
  double* s_var = met_smem + static_cast<size_t>(K) * kMetThreads;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * kMetThreads + threadIdx.x;
  const bool in = idx < B * HW;
  for (int k = 0; k < K; ++k) {
    float m = 0.f, v = 1.f;
    if (in) {
      m = means[static_cast<int64_t>(k) * B * HW + idx];
      v = expf(logvars[static_cast<int64_t>(k) * B * HW + idx]);     // float32 exp of the float32 input, as np.exp does
    }
    s_mean[k * kMetThreads + threadIdx.x] = static_cast<double>(m);
    s_var[k * kMetThreads + threadIdx.x] = static_cast<double>(v);
  }
  if (!in) return;
  const int64_t b = idx / HW, px = idx - b * HW;
  double* o = out + b * n_bins * HW + px;
  double prev = 0.0;
  for (int j = 0; j <= n_bins; ++j) {
    // np.linspace: start + j * step, the last sample is the stop value itself
    const double e = j == n_bins ? last_edge : first_edge + static_cast<double>(j) * edge_step;
    double acc = 0.0;
    for (int k = 0; k < K; ++k) {
      const double m = s_mean[k * kMetThreads + threadIdx.x], v = s_var[k * kMetThreads + threadIdx.x];
      const double z = (e - m) / v;
      acc += e < m ? exp(z) / 2 : 1 - exp(-z) / 2;                     // validate/cli.py:74-88
    }
    if (j > 0) o[static_cast<int64_t>(j - 1) * HW] = (acc - prev) / static_cast<double>(K);
    prev = acc;
  }
}

// MODE 0: kl_divergence (validate/cli.py:174-187): a += eps, g += eps, a /= sum a, g /= sum g (in place, like the
//         reference), value = sum_c g log(g / a)
// MODE 1: nll_discrete (validate/cli.py:51-70): w = g, p = a: g += eps, a += eps, g /= sum g, a /= sum a * 7,
//         value = sum_c g * -log(a)
// sums[0] += value * m, sums[1] += m (m = mask or 1)
template <int MODE>
__global__ void __launch_bounds__(128)
dist_metric_kernel(double* __restrict__ a, double* __restrict__ g, int S, int64_t B, int64_t HW,
                   const double* __restrict__ mask, double* __restrict__ value, double* __restrict__ sums) {
  __shared__ double red[4];
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  double v = 0.0, m = 0.0;
  if (idx < B * HW) {
    const int64_t b = idx / HW, px = idx - b * HW;
    double* pa = a + b * S * HW + px;
    double* pg = g + b * S * HW + px;
    const double eps = 0.00001;
    double sa = 0.0, sg = 0.0;
    for (int c = 0; c < S; ++c) {
      sa += pa[c * HW] + eps;
      sg += pg[c * HW] + eps;
    }
    if (MODE == 1) sa *= 7.0;
    for (int c = 0; c < S; ++c) {
      const double x = (pa[c * HW] + eps) / sa, y = (pg[c * HW] + eps) / sg;
      pa[c * HW] = x;
      pg[c * HW] = y;
      v += MODE == 0 ? y * log(y / x) : y * -log(x);
    }
    if (value) value[idx] = v;
    m = mask ? mask[idx] : 1.0;
  }
  const double r0 = block_sum_double(v * m, red);
  __syncthreads();
  const double r1 = block_sum_double(m, red);
  if (threadIdx.x == 0 && sums) {
    atomicAdd(sums, r0);
    atomicAdd(sums + 1, r1);
  }
}

}  // namespace mmlf

using namespace mmlf;

extern "C" int mmlf_lmm_to_discrete(const float* means, const float* logvars, int K, int64_t B, int64_t HW, int n_bins,
                                    double x_min, double x_max, double* out, void* stream) {
  MMLF_REQUIRE(means && logvars && out, "lmm_to_discrete: null buffer");
  MMLF_REQUIRE(K >= 1 && K <= 256 && n_bins >= 1 && x_max > x_min, "lmm_to_discrete: bad arguments");
  const double step = (x_max - x_min) / n_bins;
  const double first = x_min - step / 2.0, last = x_max + step / 2.0;
  const double estep = (last - first) / n_bins;                      // np.linspace(first, last, n_bins + 1)
  const size_t smem = static_cast<size_t>(2) * K * kMetThreads * sizeof(double);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(lmm_to_discrete_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    MMLF_REQUIRE(e == cudaSuccess, "lmm_to_discrete: %s", cudaGetErrorString(e));
    configured = smem;
  }
  const int64_t blocks = ceil_div64(B * HW, kMetThreads);
  lmm_to_discrete_kernel<<<static_cast<unsigned>(blocks), kMetThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      means, logvars, K, B, HW, n_bins, first, estep, last, out);
  return check_launch("lmm_to_discrete_kernel");
}

extern "C" int mmlf_kl_divergence(double* dist, double* dist_gt, int S, int64_t B, int64_t HW, const double* mask,
                                  double* value, double* sums, void* stream) {
  MMLF_REQUIRE(dist && dist_gt && S >= 1, "kl_divergence: bad arguments");
  dist_metric_kernel<0><<<static_cast<unsigned>(ceil_div64(B * HW, 128)), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      dist, dist_gt, S, B, HW, mask, value, sums);
  return check_launch("kl_divergence_kernel");
}

extern "C" int mmlf_nll_discrete(double* weights, double* posterior, int S, int64_t B, int64_t HW, const double* mask,
                                 double* value, double* sums, void* stream) {
  MMLF_REQUIRE(weights && posterior && S >= 1, "nll_discrete: bad arguments");
  dist_metric_kernel<1><<<static_cast<unsigned>(ceil_div64(B * HW, 128)), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      posterior, weights, S, B, HW, mask, value, sums);
  return check_launch("nll_discrete_kernel");
}
