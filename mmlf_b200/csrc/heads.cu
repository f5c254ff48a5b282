// Output heads, DPP target builders and the ESE reduce.  HBM-bound planar (B, C, H, W) fp32 kernels; one thread
// owns one pixel and walks the channel / member axis, so every global access is coalesced along the pixel axis.
// Reference: /root/reference/mmlf/model/feed_forward.py:270-302, utils/dl.py:109-182, model/ensamble.py:78-101.
#include "../../include/mmlf_b200.h"
#include "common.cuh"
#include "host_util.h"

namespace mmlf {

// ---------------------------------------------------------------------------- BASE / UPR head: conv(OC, OC, 2, pad 0) in fp32
__global__ void head_small_kernel(const float* __restrict__ mid, int ld_mid, int OC, const float* __restrict__ w2,
                                  const float* __restrict__ b2, int B, int H, int W, float* __restrict__ out) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t HW = static_cast<int64_t>(H) * W;
  if (idx >= B * HW) return;
  const int b = static_cast<int>(idx / HW);
  const int rem = static_cast<int>(idx - b * HW);
  const int y = rem / W, x = rem - y * W;
  const int Wp = W + 1, Hp = H + 1;
  const int64_t s00 = (static_cast<int64_t>(b) * Hp + y) * Wp + x;
  float m[4][2];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int64_t s = s00 + (t >> 1) * Wp + (t & 1);
    m[t][0] = mid[s * ld_mid];
    m[t][1] = OC > 1 ? mid[s * ld_mid + 1] : 0.f;
  }
  for (int o = 0; o < OC; ++o) {
    float acc = 0.f;
    for (int i = 0; i < OC; ++i)
#pragma unroll
      for (int t = 0; t < 4; ++t) acc = fmaf(w2[(o * OC + i) * 4 + t], m[t][i], acc);
    out[(static_cast<int64_t>(b) * OC + o) * HW + rem] = acc + b2[o];
  }
}

// gmid[s][i] = sum_o sum_t w2[o][i][t] * gout[b][o][sy - dy][sx - dx], gated by mid > 0; bf16 rows of ld_gmid channels
__global__ void head_small_bwd_data_kernel(const float* __restrict__ gout, const float* __restrict__ mid, int ld_mid,
                                           int OC, const float* __restrict__ w2, int B, int H, int W,
                                           __nv_bfloat16* __restrict__ gmid, int ld_gmid) {
  const int Wp = W + 1, Hp = H + 1;
  const int64_t n_slots = static_cast<int64_t>(B) * Hp * Wp;
  const int64_t s = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  const int b = static_cast<int>(s / (Hp * Wp));
  const int rem = static_cast<int>(s - static_cast<int64_t>(b) * Hp * Wp);
  const int sy = rem / Wp, sx = rem - sy * Wp;
  const int64_t HW = static_cast<int64_t>(H) * W;
  float g[2] = {0.f, 0.f};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int y = sy - (t >> 1), x = sx - (t & 1);
    if (y < 0 || y >= H || x < 0 || x >= W) continue;
    for (int o = 0; o < OC; ++o) {
      const float go = gout[(static_cast<int64_t>(b) * OC + o) * HW + static_cast<int64_t>(y) * W + x];
      for (int i = 0; i < OC; ++i) g[i] = fmaf(w2[(o * OC + i) * 4 + t], go, g[i]);
    }
  }
  __nv_bfloat16* dst = gmid + s * ld_gmid;
  for (int i = 0; i < ld_gmid; ++i) {
    float v = 0.f;
    if (i < OC && mid[s * ld_mid + i] > 0.f) v = g[i];
    dst[i] = __float2bfloat16_rn(v);
  }
}

// dw2[o][i][t] = sum gout[b][o][y][x] * mid[slot(b, y+dy, x+dx)][i];  db2[o] = sum gout
__global__ void __launch_bounds__(256)
head_small_bwd_param_kernel(const float* __restrict__ gout, const float* __restrict__ mid, int ld_mid, int OC, int B,
                            int H, int W, float* __restrict__ dw2, float* __restrict__ db2) {
  __shared__ float red[8][18];
  const int64_t HW = static_cast<int64_t>(H) * W;
  const int Wp = W + 1, Hp = H + 1;
  float acc[18];
#pragma unroll
  for (int j = 0; j < 18; ++j) acc[j] = 0.f;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < B * HW;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(idx / HW);
    const int rem = static_cast<int>(idx - b * HW);
    const int y = rem / W, x = rem - y * W;
    const int64_t s00 = (static_cast<int64_t>(b) * Hp + y) * Wp + x;
    for (int o = 0; o < OC; ++o) {
      const float go = gout[(static_cast<int64_t>(b) * OC + o) * HW + rem];
      acc[16 + o] += go;
      for (int i = 0; i < OC; ++i)
#pragma unroll
        for (int t = 0; t < 4; ++t)
          acc[(o * OC + i) * 4 + t] = fmaf(go, mid[(s00 + (t >> 1) * Wp + (t & 1)) * ld_mid + i], acc[(o * OC + i) * 4 + t]);
    }
  }
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 18; ++j) {
    const float v = warp_sum(acc[j]);
    if (lane == 0) red[wrp][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < 18) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    const int j = threadIdx.x;
    if (j < 16) {
      if (j < OC * OC * 4) atomicAdd(&dw2[j], v);
    } else if (j - 16 < OC) {
      atomicAdd(&db2[j - 16], v);
    }
  }
}

// ---------------------------------------------------------------------------- UPR posterior
__global__ void upr_posterior_kernel(const float* __restrict__ mean, const float* __restrict__ logvar,
                                     const float* __restrict__ bins, int steps, int64_t B, int64_t HW,
                                     float* __restrict__ post) {
  extern __shared__ float sb[];
  for (int i = threadIdx.x; i < steps; i += blockDim.x) sb[i] = bins[i];
  __syncthreads();
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= B * HW) return;
  const int64_t b = idx / HW, pix = idx - b * HW;
  const float mu = mean[idx];
  const float bb = expf(logvar[idx]);                 // "var" is used directly as the Laplace scale
  const float c = 1.0f / (2.0f * bb);
  // one reciprocal per pixel and one ex2 per bin instead of an IEEE division + expf per bin: the kernel writes 432 B per
  // pixel and was issue bound, not HBM bound.  |x - mu| * (1 / b) differs from |x - mu| / b by <= 1 ulp of the exponent
  // argument, i.e. by |arg| * 1.2e-7 relative in the result (arg <= ~50 before the value underflows the 2e-5 tolerance).
  const float nrb = -1.4426950408889634f / bb;        // -log2(e) / b
  float* o = post + b * steps * HW + pix;
#pragma unroll 4
  for (int j = 0; j < steps; ++j) __stcs(o + j * HW, c * exp2f(fabsf(sb[j] - mu) * nrb));
}

// ---------------------------------------------------------------------------- DPP head
constexpr int kHeadUnroll = 12;

__global__ void dpp_head_kernel(const float* __restrict__ scores, const float* __restrict__ bins_t,
                                const float* __restrict__ bins_n, int steps, int64_t B, int64_t HW,
                                float* __restrict__ one_hot, float* __restrict__ post, float* __restrict__ mean,
                                float* __restrict__ logvar) {
  extern __shared__ float sb[];                        // bins_t | bins_n
  for (int i = threadIdx.x; i < steps; i += blockDim.x) {
    sb[i] = bins_t[i];
    sb[steps + i] = bins_n[i];
  }
  __syncthreads();
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= B * HW) return;
  const int64_t b = idx / HW, pix = idx - b * HW;
  const float* s = scores + b * steps * HW + pix;
  // Every loop handles kHeadUnroll channel planes per iteration with the loads issued first (one 4-byte load per plane
  // and thread).  Passes 2 and 3 re-read the scores from L2.  The kernel is issue bound: running pointers, exp through
  // ex2.approx (__expf, relative error 2^-21; the posterior is checked to 2e-5) and one reciprocal of z per pixel.
  float mx = -INFINITY, z = 0.f;
  const float* sp = s;
  for (int c0 = 0; c0 < steps; c0 += kHeadUnroll) {
    float v[kHeadUnroll];
#pragma unroll
    for (int u = 0; u < kHeadUnroll; ++u) {
      v[u] = c0 + u < steps ? __ldg(sp) : -INFINITY;
      sp += HW;
    }
#pragma unroll
    for (int u = 0; u < kHeadUnroll; ++u) {
      if (c0 + u < steps) {
        mx = fmaxf(mx, v[u]);
        z += __expf(v[u]);                               // unstabilised, as the reference
      }
    }
  }
  const float rz = 1.f / z;
  float mu = 0.f;
  sp = s;
  float* ohp = one_hot ? one_hot + b * steps * HW + pix : nullptr;
  float* pp = post ? post + b * steps * HW + pix : nullptr;
  for (int c0 = 0; c0 < steps; c0 += kHeadUnroll) {
    float v[kHeadUnroll];
#pragma unroll
    for (int u = 0; u < kHeadUnroll; ++u) {
      v[u] = c0 + u < steps ? __ldg(sp) : 0.f;
      sp += HW;
    }
#pragma unroll
    for (int u = 0; u < kHeadUnroll; ++u) {
      const int c = c0 + u;
      if (c < steps) {
        const float oh = (v[u] == mx) ? 1.f : 0.f;       // ties give a multi-hot vector
        mu += sb[c] * oh;
        const float pc = __expf(v[u]) * rz;
        if (ohp) __stcs(ohp, oh);
        if (pp) __stcs(pp, pc);
        v[u] = pc;
      }
      if (ohp) ohp += HW;
      if (pp) pp += HW;
    }
  }
  float acc = 0.f;
  sp = s;
  for (int c0 = 0; c0 < steps; c0 += kHeadUnroll) {
    float v[kHeadUnroll];
#pragma unroll
    for (int u = 0; u < kHeadUnroll; ++u) {
      v[u] = c0 + u < steps ? __ldg(sp) : 0.f;
      sp += HW;
    }
#pragma unroll
    for (int u = 0; u < kHeadUnroll; ++u) {
      const int c = c0 + u;
      if (c < steps) {
        const float d = sb[steps + c] - mu;
        acc = fmaf(d * d, __expf(v[u]) * rz, acc);
      }
    }
  }
  mean[idx] = mu;
  logvar[idx] = logf(acc);
}

// Shared-memory variant (HW % 128 == 0, 16-byte aligned scores): the block's [steps][128 pixels] score tile is fetched
// ONCE with 16-byte cp.async (no registers, the whole 55 KB tile in flight per block) and all three passes run from
// shared memory.  The single-pass-per-plane kernel above re-read the scores for passes 2 and 3 and, at 255 MB per launch
// against 126 MB of L2, most of those re-reads went to DRAM (ncu: 714 MB read for 257 MB algorithmic).
constexpr int kHeadTile = 128;

__global__ void __launch_bounds__(kHeadTile)
dpp_head_smem_kernel(const float* __restrict__ scores, const float* __restrict__ bins_t, const float* __restrict__ bins_n,
                     int steps, int64_t HW, float* __restrict__ one_hot, float* __restrict__ post,
                     float* __restrict__ mean, float* __restrict__ logvar) {
  extern __shared__ __align__(16) float hs[];         // tile [steps][128] | bins_t | bins_n
  float* tile = hs;
  float* sb = hs + static_cast<size_t>(steps) * kHeadTile;
  const int64_t idx0 = static_cast<int64_t>(blockIdx.x) * kHeadTile;     // first pixel of the block (never straddles b)
  const int64_t b = idx0 / HW, pix0 = idx0 - b * HW;
  const float* s = scores + b * steps * HW + pix0;
  for (int i = threadIdx.x; i < steps * (kHeadTile / 4); i += kHeadTile) {
    const int c = i / (kHeadTile / 4), q = i - c * (kHeadTile / 4);
    const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(tile + c * kHeadTile + 4 * q));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(s + static_cast<int64_t>(c) * HW + 4 * q) : "memory");
  }
  for (int i = threadIdx.x; i < steps; i += kHeadTile) {
    sb[i] = bins_t[i];
    sb[steps + i] = bins_n[i];
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  float* col = tile + threadIdx.x;                    // this thread's pixel: conflict-free column walk
  float mx = -INFINITY, z = 0.f;
#pragma unroll 4
  for (int c = 0; c < steps; ++c) {
    const float v = col[c * kHeadTile];
    mx = fmaxf(mx, v);
    z += __expf(v);                                   // unstabilised, as the reference
  }
  const float rz = 1.f / z;
  const int64_t o0 = b * steps * HW + pix0 + threadIdx.x;
  float* ohp = one_hot ? one_hot + o0 : nullptr;
  float* pp = post ? post + o0 : nullptr;
  float mu = 0.f;
#pragma unroll 4
  for (int c = 0; c < steps; ++c) {
    const float v = col[c * kHeadTile];
    const float oh = (v == mx) ? 1.f : 0.f;           // ties give a multi-hot vector
    mu += sb[c] * oh;
    const float pc = __expf(v) * rz;
    col[c * kHeadTile] = pc;                          // pass 3 reads the posterior back
    if (ohp) __stcs(ohp + static_cast<int64_t>(c) * HW, oh);
    if (pp) __stcs(pp + static_cast<int64_t>(c) * HW, pc);
  }
  float acc = 0.f;
#pragma unroll 4
  for (int c = 0; c < steps; ++c) {
    const float d = sb[steps + c] - mu;
    acc = fmaf(d * d, col[c * kHeadTile], acc);
  }
  mean[idx0 + threadIdx.x] = mu;
  logvar[idx0 + threadIdx.x] = logf(acc);
}

// ---------------------------------------------------------------------------- DPP targets
__global__ void reg_to_class_kernel(const float* __restrict__ gt, const float* __restrict__ bins, int steps,
                                    float half_step, int64_t B, int64_t HW, float* __restrict__ out) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= B * HW) return;
  const int64_t b = idx / HW, pix = idx - b * HW;
  const float g = gt[idx];
  float* o = out + b * steps * HW + pix;
  for (int c = 0; c < steps; ++c) o[c * HW] = fabsf(__ldg(bins + c) - g) < half_step ? 1.f : 0.f;
}

__global__ void mpi_to_weights_kernel(const float* __restrict__ mpi, int K, const float* __restrict__ bins, int steps,
                                      float half_step, int64_t B, int64_t HW, float* __restrict__ out) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= B * HW) return;
  const int64_t b = idx / HW, pix = idx - b * HW;
  float w[12], d[12];
  for (int k = 0; k < K; ++k) {
    w[k] = mpi[((b * K + k) * 5 + 3) * HW + pix];
    d[k] = mpi[((b * K + k) * 5 + 4) * HW + pix];
  }
  float* o = out + b * steps * HW + pix;
  for (int c = 0; c < steps; ++c) {
    const float bin = __ldg(bins + c);
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc += (fabsf(bin - d[k]) < half_step ? 1.f : 0.f) * w[k];
    o[c * HW] = acc;
  }
}

// ---------------------------------------------------------------------------- ESE reduce
constexpr int kEseThreads = 64;
__global__ void __launch_bounds__(kEseThreads)
ese_reduce_kernel(const float* __restrict__ means, const float* __restrict__ logvars, const float* __restrict__ disp,
                  int K, int64_t B, int64_t HW, float* __restrict__ mean, float* __restrict__ logvar,
                  float* __restrict__ post) {
  extern __shared__ float sm[];                        // m[K][T] | b[K][T] | c[K][T] | disp[K]
  float* sm_m = sm;
  float* sm_b = sm + K * kEseThreads;
  float* sm_c = sm_b + K * kEseThreads;
  float* sm_d = sm_c + K * kEseThreads;
  for (int i = threadIdx.x; i < K; i += kEseThreads) sm_d[i] = disp[i];
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * kEseThreads + threadIdx.x;
  const bool live = idx < B * HW;
  const int t = threadIdx.x;
  if (live) {
    float best = 0.f, best_m = 0.f;
    for (int i = 0; i < K; ++i) {
      const float lv = __ldg(logvars + i * B * HW + idx);
      const float m = __ldg(means + i * B * HW + idx);
      if (i == 0 || lv < best) {                       // first minimum wins, like torch.min
        best = lv;
        best_m = m;
      }
      const float bb = expf(lv);
      sm_m[i * kEseThreads + t] = m;
      sm_b[i * kEseThreads + t] = -1.4426950408889634f / bb;     // -log2(e) / b: one reciprocal per member, not per pair
      sm_c[i * kEseThreads + t] = 1.0f / (2.0f * bb);
    }
    mean[idx] = best_m;
    logvar[idx] = best;
  }
  __syncthreads();
  if (!live) return;
  const int64_t b = idx / HW, pix = idx - b * HW;
  const float kf = static_cast<float>(K);
  for (int j = 0; j < K; ++j) {
    const float x = sm_d[j];
    float acc = 0.f;
    // K x K Laplace evaluations per pixel (4900 for the 70-member ensemble): one add, one multiply, one ex2 and one FMA each
    // (was an IEEE division + expf: 2.5 ms per 512 x 512 light field, 11 % of the SFU rate).  The exponent argument
    // differs from -|x - m| / b by <= 2 ulp, i.e. by |arg| * 2.4e-7 relative in the term.
#pragma unroll 2
    for (int i = 0; i < K; ++i)
      acc = fmaf(sm_c[i * kEseThreads + t], exp2f(fabsf(x - sm_m[i * kEseThreads + t]) * sm_b[i * kEseThreads + t]), acc);
    __stcs(post + (b * K + j) * HW + pix, acc / kf);
  }
}

}  // namespace mmlf

using namespace mmlf;

static inline unsigned blocks_for(int64_t n, int t) { return static_cast<unsigned>(ceil_div64(n, t)); }

extern "C" int mmlf_head_small(const float* mid, int ld_mid, int OC, const float* w2, const float* b2, int B, int H,
                               int W, float* out, void* stream) {
  MMLF_REQUIRE(mid && w2 && b2 && out, "head_small: null buffer");
  MMLF_REQUIRE(OC == 1 || OC == 2, "head_small: OC must be 1 or 2 (got %d)", OC);
  const int64_t n = static_cast<int64_t>(B) * H * W;
  head_small_kernel<<<blocks_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(mid, ld_mid, OC, w2, b2, B, H, W, out);
  return check_launch("head_small");
}

extern "C" int mmlf_head_small_bwd(const float* gout, const float* mid, int ld_mid, int OC, const float* w2, int B,
                                   int H, int W, void* gmid, int ld_gmid, float* dw2, float* db2, void* stream) {
  MMLF_REQUIRE(gout && mid && w2 && gmid && dw2 && db2, "head_small_bwd: null buffer");
  MMLF_REQUIRE(OC == 1 || OC == 2, "head_small_bwd: OC must be 1 or 2");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t n_slots = static_cast<int64_t>(B) * (H + 1) * (W + 1);
  head_small_bwd_data_kernel<<<blocks_for(n_slots, 256), 256, 0, st>>>(gout, mid, ld_mid, OC, w2, B, H, W,
                                                                     reinterpret_cast<__nv_bfloat16*>(gmid), ld_gmid);
  if (int rc = check_launch("head_small_bwd_data")) return rc;
  const int64_t n = static_cast<int64_t>(B) * H * W;
  int grid = static_cast<int>(ceil_div64(n, 256 * 8));
  if (grid > sm_count() * 4) grid = sm_count() * 4;
  if (grid < 1) grid = 1;
  head_small_bwd_param_kernel<<<grid, 256, 0, st>>>(gout, mid, ld_mid, OC, B, H, W, dw2, db2);
  return check_launch("head_small_bwd_param");
}

extern "C" int mmlf_upr_posterior(const float* mean, const float* logvar, const float* bins, int steps, int64_t B,
                                  int64_t HW, float* posterior, void* stream) {
  MMLF_REQUIRE(mean && logvar && bins && posterior, "upr_posterior: null buffer");
  upr_posterior_kernel<<<blocks_for(B * HW, 256), 256, steps * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      mean, logvar, bins, steps, B, HW, posterior);
  return check_launch("upr_posterior");
}

extern "C" int mmlf_dpp_head(const float* scores, const float* bins_t, const float* bins_n, int steps, int64_t B,
                             int64_t HW, float* one_hot, float* posterior, float* mean, float* logvar, void* stream) {
  MMLF_REQUIRE(scores && bins_t && bins_n && mean && logvar, "dpp_head: null buffer");
  const size_t tile_smem = (static_cast<size_t>(steps) * kHeadTile + 2 * steps) * sizeof(float);
  if (HW % kHeadTile == 0 && tile_smem <= 200 * 1024 && reinterpret_cast<uintptr_t>(scores) % 16 == 0) {
    static size_t configured = 0;
    if (tile_smem > 48 * 1024 && tile_smem > configured) {
      cudaError_t e = cudaFuncSetAttribute(dpp_head_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      MMLF_REQUIRE(e == cudaSuccess, "dpp_head: %s", cudaGetErrorString(e));
      configured = 200 * 1024;
    }
    dpp_head_smem_kernel<<<static_cast<unsigned>(B * HW / kHeadTile), kHeadTile, tile_smem, static_cast<cudaStream_t>(stream)>>>(
        scores, bins_t, bins_n, steps, HW, one_hot, posterior, mean, logvar);
    return check_launch("dpp_head_smem");
  }
  dpp_head_kernel<<<blocks_for(B * HW, 128), 128, 2 * steps * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      scores, bins_t, bins_n, steps, B, HW, one_hot, posterior, mean, logvar);
  return check_launch("dpp_head");
}

extern "C" int mmlf_reg_to_class(const float* gt, const float* bins_t, int steps, double half_step, int64_t B,
                                 int64_t HW, float* out, void* stream) {
  MMLF_REQUIRE(gt && bins_t && out, "reg_to_class: null buffer");
  reg_to_class_kernel<<<blocks_for(B * HW, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      gt, bins_t, steps, static_cast<float>(half_step), B, HW, out);
  return check_launch("reg_to_class");
}

extern "C" int mmlf_mpi_to_weights(const float* mpi, int K, const float* bins_t, int steps, double half_step,
                                   int64_t B, int64_t HW, float* out, void* stream) {
  MMLF_REQUIRE(mpi && bins_t && out, "mpi_to_weights: null buffer");
  MMLF_REQUIRE(K >= 1 && K <= 12, "mpi_to_weights: K must be in [1, 12] (hci4d.py:221-222 caps the planes at 12)");
  mpi_to_weights_kernel<<<blocks_for(B * HW, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      mpi, K, bins_t, steps, static_cast<float>(half_step), B, HW, out);
  return check_launch("mpi_to_weights");
}

extern "C" int mmlf_ese_reduce(const float* means, const float* logvars, const float* disp, int K, int64_t B,
                               int64_t HW, float* mean, float* logvar, float* posterior, void* stream) {
  MMLF_REQUIRE(means && logvars && disp && mean && logvar && posterior, "ese_reduce: null buffer");
  MMLF_REQUIRE(K >= 1 && K <= 256, "ese_reduce: K must be in [1, 256]");
  const size_t smem = (static_cast<size_t>(3) * K * kEseThreads + K) * sizeof(float);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(ese_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    MMLF_REQUIRE(e == cudaSuccess, "ese_reduce: %s", cudaGetErrorString(e));
    configured = smem;
  }
  ese_reduce_kernel<<<blocks_for(B * HW, kEseThreads), kEseThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      means, logvars, disp, K, B, HW, mean, logvar, posterior);
  return check_launch("ese_reduce");
}
