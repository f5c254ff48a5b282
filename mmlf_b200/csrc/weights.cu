// Weight packing for the implicit-GEMM convolution and the inverse mapping for weight gradients.
// The spatial variants replace the permute/flip plumbing of /root/reference/mmlf/model/feed_forward.py:236-256:
// all four streams stay in native layout and use per-stream views of the shared weights.
#include "../../include/mmlf_b200.h"
#include "common.cuh"
#include "host_util.h"

namespace mmlf {

// canonical tap (a, b) of w[., ., a, b] that the effective tap (p, q) of a stream reads
__host__ __device__ __forceinline__ void eff_to_canonical(int spatial, int p, int q, int& a, int& b) {
  if (spatial == 0) { a = p; b = q; }            // v / d streams
  else if (spatial == 1) { a = q; b = p; }       // h stream: transposed
  else { a = q; b = 1 - p; }                     // i stream: flip(w, -1).permute(0, 1, 3, 2)
}

// padded channel index -> real channel index, or -1 for a padding channel
__host__ __device__ __forceinline__ int real_channel(int c, int groups, int group_real, int group_pad) {
  const int g = c / group_pad, cc = c - g * group_pad;
  if (g >= groups || cc >= group_real) return -1;
  return g * group_real + cc;
}

__device__ __forceinline__ void pack_conv_weight_elem(int64_t idx, const float* __restrict__ w, int cout, int cin,
                                                      int spatial, int dgrad, int groups, int group_real, int group_pad,
                                                      uint16_t* __restrict__ out, int n_pad, int n_kc, int dtype,
                                                      int split, float wscale) {
  // split: three K blocks per tap, [w_hi | w_lo | w_hi]
  const int terms = split ? 3 : 1;
  const int k_total = 4 * terms * n_kc * 64;
  if (idx >= static_cast<int64_t>(n_pad) * k_total) return;
  const int row = static_cast<int>(idx / k_total);
  const int col = static_cast<int>(idx - static_cast<int64_t>(row) * k_total);
  const int tap = col / (terms * n_kc * 64);
  const int ct = col - tap * terms * n_kc * 64;
  const int term = ct / (n_kc * 64), c = ct - term * n_kc * 64;
  int p = tap >> 1, q = tap & 1;
  float v = 0.f;
  int n, ci;
  if (!dgrad) {
    n = row < cout ? row : -1;
    ci = real_channel(c, groups, group_real, group_pad);
  } else {
    // data gradient: GEMM rows = input channels, K = output channels, taps rotated by 180 degrees
    ci = real_channel(row, groups, group_real, group_pad);
    n = c < cout ? c : -1;
    p = 1 - p;
    q = 1 - q;
  }
  if (n >= 0 && ci >= 0 && ci < cin) {
    int a, b;
    eff_to_canonical(spatial, p, q, a, b);
    v = w[((static_cast<int64_t>(n) * cin + ci) * 2 + a) * 2 + b];
  }
  v *= wscale;                                                      // power of two: exact
  if (split && term == 1) v -= from16(to16(v, kFP16), kFP16);      // residual of the fp16 rounding
  out[idx] = to16(v, split ? kFP16 : dtype);
}

__global__ void pack_conv_weight_kernel(const float* __restrict__ w, int cout, int cin, int spatial, int dgrad,
                                        int groups, int group_real, int group_pad, uint16_t* __restrict__ out,
                                        int n_pad, int n_kc, int dtype, int split, float wscale) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  pack_conv_weight_elem(idx, w, cout, cin, spatial, dgrad, groups, group_real, group_pad, out, n_pad, n_kc, dtype, split, wscale);
}

// blockIdx.y = job, blockIdx.x strides over the job's elements (the grid is sized for a fraction of the largest job)
__global__ void __launch_bounds__(256)
pack_conv_weights_batch_kernel(const mmlf_pack_job* __restrict__ jobs) {
  const mmlf_pack_job j = jobs[blockIdx.y];
  const int n_kc = (j.cin_pad + 63) / 64;
  const int64_t total = static_cast<int64_t>(j.n_pad) * 4 * (j.split ? 3 : 1) * n_kc * 64;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x)
    pack_conv_weight_elem(idx, j.w, j.cout, j.cin, j.spatial, j.dgrad, j.in_groups, j.group_real, j.group_pad,
                          reinterpret_cast<uint16_t*>(j.out), j.n_pad, n_kc, j.dtype, j.split, j.weight_scale);
  if (j.bias_pad && blockIdx.x == 0)
    for (int i = threadIdx.x; i < j.n_pad; i += blockDim.x) j.bias_pad[i] = (j.bias && i < j.cout) ? j.bias[i] : 0.f;
}

__global__ void unpack_conv_wgrad_kernel(const float* __restrict__ dwp, int n_pad, int cin_pad, int cout, int cin,
                                         int spatial, int groups, int group_real, int group_pad,
                                         float* __restrict__ dw, int accumulate) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<int64_t>(cout) * cin * 4) return;
  const int b = idx & 1, a = (idx >> 1) & 1;
  const int ci = static_cast<int>((idx >> 2) % cin), n = static_cast<int>((idx >> 2) / cin);
  // inverse of eff_to_canonical
  int p, q;
  if (spatial == 0) { p = a; q = b; }
  else if (spatial == 1) { p = b; q = a; }
  else { p = 1 - b; q = a; }
  const int g = ci / group_real, cc = ci - g * group_real;
  const int c = g * group_pad + cc;
  const float v = dwp[(static_cast<int64_t>(n) * 4 + (p * 2 + q)) * cin_pad + c];
  dw[idx] = accumulate ? dw[idx] + v : v;
}

}  // namespace mmlf

using namespace mmlf;

extern "C" int mmlf_pack_conv_weight(const float* w, int cout, int cin, int spatial, int dgrad, int in_groups,
                                     int group_real, int group_pad, void* out, int n_pad, int cin_pad, int dtype,
                                     void* stream) {
  MMLF_REQUIRE(w && out, "pack_conv_weight: null buffer");
  MMLF_REQUIRE(dtype == 0 || dtype == 1, "pack_conv_weight: dtype must be 0 (bf16) or 1 (fp16)");
  MMLF_REQUIRE(spatial >= 0 && spatial <= 2, "pack_conv_weight: spatial must be 0..2");
  MMLF_REQUIRE(in_groups >= 1 && group_real >= 1 && group_pad >= group_real, "pack_conv_weight: bad channel groups");
  MMLF_REQUIRE(in_groups * group_real == cin, "pack_conv_weight: groups (%d x %d) do not cover cin %d", in_groups, group_real, cin);
  MMLF_REQUIRE(n_pad % 16 == 0 && cin_pad % 16 == 0, "pack_conv_weight: pads must be multiples of 16");
  if (!dgrad)
    MMLF_REQUIRE(n_pad >= cout && cin_pad >= in_groups * group_pad, "pack_conv_weight: pads too small");
  else
    MMLF_REQUIRE(n_pad >= in_groups * group_pad && cin_pad >= cout, "pack_conv_weight(dgrad): pads too small");
  const int n_kc = ceil_div(cin_pad, 64);
  const int64_t total = static_cast<int64_t>(n_pad) * 4 * n_kc * 64;
  pack_conv_weight_kernel<<<static_cast<unsigned>(ceil_div64(total, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, cout, cin, spatial, dgrad, in_groups, group_real, group_pad, reinterpret_cast<uint16_t*>(out), n_pad, n_kc, dtype, 0, 1.0f);
  return check_launch("pack_conv_weight_kernel");
}

extern "C" int mmlf_pack_conv_weight_split(const float* w, int cout, int cin, int spatial, int in_groups, int group_real,
                                           int group_pad, void* out, int n_pad, int cin_pad, float weight_scale, void* stream) {
  MMLF_REQUIRE(w && out, "pack_conv_weight_split: null buffer");
  MMLF_REQUIRE(weight_scale > 0.f, "pack_conv_weight_split: weight_scale must be positive");
  MMLF_REQUIRE(spatial >= 0 && spatial <= 2, "pack_conv_weight_split: spatial must be 0..2");
  MMLF_REQUIRE(in_groups >= 1 && group_real >= 1 && group_pad >= group_real && in_groups * group_real == cin,
               "pack_conv_weight_split: bad channel groups");
  MMLF_REQUIRE(n_pad % 16 == 0 && cin_pad % 16 == 0 && n_pad >= cout && cin_pad >= in_groups * group_pad,
               "pack_conv_weight_split: bad pads");
  const int n_kc = ceil_div(cin_pad, 64);
  const int64_t total = static_cast<int64_t>(n_pad) * 4 * 3 * n_kc * 64;
  pack_conv_weight_kernel<<<static_cast<unsigned>(ceil_div64(total, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, cout, cin, spatial, 0, in_groups, group_real, group_pad, reinterpret_cast<uint16_t*>(out), n_pad, n_kc, kFP16, 1, weight_scale);
  return check_launch("pack_conv_weight_kernel(split)");
}

extern "C" int mmlf_pack_conv_weights_batch(const mmlf_pack_job* jobs, int n_jobs, int64_t max_elems, void* stream) {
  static_assert(sizeof(mmlf_pack_job) == 80, "mmlf_pack_job layout (mirrored by mmlf_b200/engine.py::PackJob)");
  MMLF_REQUIRE(jobs && n_jobs >= 1 && n_jobs <= 65535 && max_elems >= 1, "pack_conv_weights_batch: bad arguments");
  // 16 elements per thread for the largest job: ~90 blocks per 280 x 280 layer, one wave for the whole network
  const unsigned bx = static_cast<unsigned>(ceil_div64(max_elems, 256 * 16));
  pack_conv_weights_batch_kernel<<<dim3(bx, static_cast<unsigned>(n_jobs)), 256, 0, static_cast<cudaStream_t>(stream)>>>(jobs);
  return check_launch("pack_conv_weights_batch_kernel");
}

extern "C" int mmlf_unpack_conv_wgrad(const float* dw_packed, int n_pad, int cin_pad, int cout, int cin, int spatial,
                                      int in_groups, int group_real, int group_pad, float* dw, int accumulate,
                                      void* stream) {
  MMLF_REQUIRE(dw_packed && dw, "unpack_conv_wgrad: null buffer");
  MMLF_REQUIRE(in_groups * group_real == cin && n_pad >= cout && cin_pad >= in_groups * group_pad,
               "unpack_conv_wgrad: inconsistent channel layout");
  const int64_t total = static_cast<int64_t>(cout) * cin * 4;
  unpack_conv_wgrad_kernel<<<static_cast<unsigned>(ceil_div64(total, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dw_packed, n_pad, cin_pad, cout, cin, spatial, in_groups, group_real, group_pad, dw, accumulate);
  return check_launch("unpack_conv_wgrad_kernel");
}
