// BatchNorm (training and eval), ReLU backward and column sums on the 16-bit slot layout (fp16 activations,
// bf16 gradients; dtype codes as in common.cuh).
// Replaces nn.BatchNorm2d / nn.ReLU of /root/reference/mmlf/model/feed_forward.py:134-135 and their autograd.
// All kernels are HBM-bound: 128-bit loads/stores, one thread = 8 consecutive channels of one slot, so a warp
// reads contiguous runs of the channel-last rows.
#include "../../include/mmlf_b200.h"
#include "common.cuh"
#include <stdlib.h>
#include "host_util.h"

namespace mmlf {

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8], int dt) {
  unpack16x2(u.x, dt, f[0], f[1]);
  unpack16x2(u.y, dt, f[2], f[3]);
  unpack16x2(u.z, dt, f[4], f[5]);
  unpack16x2(u.w, dt, f[6], f[7]);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8], int dt) {
  uint4 u;
  u.x = pack16x2(f[0], f[1], dt); u.y = pack16x2(f[2], f[3], dt);
  u.z = pack16x2(f[4], f[5], dt); u.w = pack16x2(f[6], f[7], dt);
  return u;
}
__device__ __forceinline__ uint4 ld8(const void* base, int64_t slot, int ld, int c) {
  return __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + slot * ld + c));
}
__device__ __forceinline__ void st8(void* base, int64_t slot, int ld, int c, const uint4& v) {
  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + slot * ld + c) = v;
}
__device__ __forceinline__ bool slot_valid(int64_t s, int Hp, int Wp) {
  const uint32_t rem = static_cast<uint32_t>(s) % static_cast<uint32_t>(Hp * Wp);   // n_slots < 2^31
  const uint32_t sy = rem / static_cast<uint32_t>(Wp), sx = rem - sy * Wp;
  return sy >= 1 && sx >= 1;
}

// ---------------------------------------------------------------------------- column reductions
// MODE 0: sum x, sum x^2 (bn_stats)      MODE 1: sum x (colsum)
// MODE 2: g = dy * (z * scale + shift > 0), xhat = (z - mean) * invstd: sum g, sum g * xhat (bn_bwd_reduce); the ReLU
//         mask is recomputed from z exactly as the forward pass computed y = relu(fma(z, scale, shift))
constexpr int kRedThreads = 256;
constexpr int kUnroll = 4;

template <int MODE>
__global__ void __launch_bounds__(kRedThreads)
col_reduce_kernel(const void* __restrict__ x, int ld_x, const float* __restrict__ scale,
                  const float* __restrict__ shift, const void* __restrict__ z, int ld_z, const float* __restrict__ mean,
                  const float* __restrict__ invstd, int C, int64_t n_slots, double* __restrict__ sums,
                  float* __restrict__ fsum, int dt_x, int dt_yz) {
  extern __shared__ float red[];                    // [lanes][groups * 16]
  const int groups = C >> 3;
  const int lanes = kRedThreads / groups;           // slot lanes per block
  const int g = threadIdx.x % groups, sl = threadIdx.x / groups;
  float a0[8], a1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a0[j] = a1[j] = 0.f;
  float mu[8], is[8], sc[8], sh[8];
  if (MODE == 2) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mu[j] = mean[g * 8 + j]; is[j] = invstd[g * 8 + j]; sc[j] = scale[g * 8 + j]; sh[j] = shift[g * 8 + j];
    }
  }
  if (sl < lanes) {
    // kUnroll slots per iteration with all loads issued first: one slot's two 16-byte loads per thread in flight kept
    // the kernel at 58 % of the HBM rate (latency bound)
    const int64_t step = static_cast<int64_t>(gridDim.x) * lanes;
    for (int64_t s = static_cast<int64_t>(blockIdx.x) * lanes + sl; s < n_slots; s += kUnroll * step) {
      uint4 xv[kUnroll], zv[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int64_t su = s + u * step;
        xv[u] = make_uint4(0u, 0u, 0u, 0u);      // zero bit patterns contribute nothing to any of the sums
        zv[u] = xv[u];
        if (su < n_slots) {
          xv[u] = ld8(x, su, ld_x, g * 8);
          if (MODE == 2) zv[u] = ld8(z, su, ld_z, g * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        float v[8];
        unpack8(xv[u], v, dt_x);
        if (MODE == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { a0[j] += v[j]; a1[j] = fmaf(v[j], v[j], a1[j]); }
        } else if (MODE == 1) {
#pragma unroll
          for (int j = 0; j < 8; ++j) a0[j] += v[j];
        } else {
          float zz[8];
          unpack8(zv[u], zz, dt_yz);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float gg = fmaf(zz[j], sc[j], sh[j]) > 0.f ? v[j] : 0.f;
            a0[j] += gg;
            a1[j] = fmaf(gg, (zz[j] - mu[j]) * is[j], a1[j]);
          }
        }
      }
    }
  }
  // block reduction over the slot lanes
  const int stride = groups * 16;
  if (sl < lanes) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      red[sl * stride + g * 16 + j] = a0[j];
      red[sl * stride + g * 16 + 8 + j] = a1[j];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < stride; i += kRedThreads) {
    double acc = 0.0;
    for (int l = 0; l < lanes; ++l) acc += static_cast<double>(red[l * stride + i]);
    const int gg = i / 16, j = i % 16;
    const int c = gg * 8 + (j & 7);
    if (MODE == 1) {
      if (j < 8) atomicAdd(&fsum[c], static_cast<float>(acc));
    } else {
      atomicAdd(&sums[(j >> 3) * C + c], acc);
    }
  }
}

// ---------------------------------------------------------------------------- finalize (training)
__global__ void bn_finalize_kernel(const double* __restrict__ sums, int C_real, int C, int64_t count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ rmean, float* __restrict__ rvar, int64_t* __restrict__ nbt,
                                   float momentum, float eps, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && nbt) *nbt += 1;
  if (c >= C) return;
  if (c >= C_real) {                                // padding channels stay exactly zero
    scale[c] = 0.f; shift[c] = 0.f; save_mean[c] = 0.f; save_invstd[c] = 0.f;
    return;
  }
  const double n = static_cast<double>(count);
  const double mean = sums[c] / n;
  double var = sums[C + c] / n - mean * mean;       // biased, used to normalise
  if (var < 0.0) var = 0.0;
  const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float sc = gamma[c] * invstd;
  scale[c] = sc;
  shift[c] = beta[c] - static_cast<float>(mean) * sc;
  save_mean[c] = static_cast<float>(mean);
  save_invstd[c] = invstd;
  if (rmean) {
    const double unbiased = count > 1 ? var * n / (n - 1.0) : var;
    rmean[c] = static_cast<float>((1.0 - momentum) * rmean[c] + momentum * mean);
    rvar[c] = static_cast<float>((1.0 - momentum) * rvar[c] + momentum * unbiased);
  }
}

__global__ void bn_fold_eval_kernel(int C_real, int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ rmean, const float* __restrict__ rvar,
                                    const float* __restrict__ conv_bias, float eps, float* __restrict__ scale,
                                    float* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (c >= C_real) { scale[c] = 0.f; shift[c] = 0.f; return; }
  const float invstd = static_cast<float>(1.0 / sqrt(static_cast<double>(rvar[c]) + static_cast<double>(eps)));
  const float sc = gamma[c] * invstd;
  scale[c] = sc;
  shift[c] = ((conv_bias ? conv_bias[c] : 0.f) - rmean[c]) * sc + beta[c];
}

// ---------------------------------------------------------------------------- elementwise passes
// MODE 0: y = relu(z * scale + shift), halo -> 0; optional second copy of y in another 16-bit format (bn_apply_relu)
// MODE 1: dz = dy * (y > 0)                                          (relu_bwd)
// MODE 2: dz = gamma*invstd*(g - sg/n - xhat*sgx/n), halo -> 0       (bn_bwd_apply, train)
// MODE 3: dz = g * gamma * invstd                                    (bn_bwd_apply, eval-mode BN)
//         in modes 2/3 g = dy * (z * scale + shift > 0) and, if csum is given, csum[c] += sum_slots dz (as stored):
//         the bias gradient of the convolution in front of the BatchNorm
// MODE 4: 16-bit format conversion                                   (convert16)
// block = (channel groups) x (slot lanes); a thread owns 8 consecutive channels, keeps their per-channel constants
// in registers and walks the slots with a grid stride, so the per-slot work is 2-4 128-bit memory operations.
struct SlotDiv {   // s -> (sy >= 1 && sx >= 1) without integer division when the slot index fits a float exactly
  float rcp_img, rcp_w;
  uint32_t per_img, Wp;
  int exact;
};
__device__ __forceinline__ uint32_t fast_div(uint32_t x, uint32_t d, float rcp) {
  uint32_t q = static_cast<uint32_t>(__uint2float_rz(x) * rcp);
  int32_t r = static_cast<int32_t>(x - q * d);
  if (r < 0) { --q; r += d; }
  if (r >= static_cast<int32_t>(d)) ++q;
  return q;
}
__device__ __forceinline__ bool slot_valid_fast(int64_t s, const SlotDiv& d) {
  const uint32_t x = static_cast<uint32_t>(s);
  uint32_t rem, sy;
  if (d.exact) {
    rem = x - fast_div(x, d.per_img, d.rcp_img) * d.per_img;
    sy = fast_div(rem, d.Wp, d.rcp_w);
  } else {
    rem = x % d.per_img;
    sy = rem / d.Wp;
  }
  const uint32_t sx = rem - sy * d.Wp;
  return sy >= 1 && sx >= 1;
}

template <int MODE>
__global__ void __launch_bounds__(256)
slot_map_kernel(const void* __restrict__ a, int ld_a, const void* __restrict__ y, int ld_y,
                const void* __restrict__ z, int ld_z, const float* __restrict__ p0, const float* __restrict__ p1,
                const float* __restrict__ p2, const float* __restrict__ fsums, const float* __restrict__ scale,
                const float* __restrict__ shift, int C, SlotDiv dv, int64_t n_slots, void* __restrict__ out, int ld_out,
                void* __restrict__ out2, int ld_out2, int dt_out2, float* __restrict__ csum, int dt_a, int dt_yz) {
  extern __shared__ float map_red[];               // [lanes][groups * 8], only used with csum
  const int groups = C >> 3;
  const int lanes = blockDim.x / groups;
  const int g = threadIdx.x % groups, sl = threadIdx.x / groups;
  const bool active = sl < lanes;
  if (!active && !((MODE == 2 || MODE == 3) && csum)) return;
  const int c = g * 8;
  float cs[8], msc[8], msh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) cs[j] = msc[j] = msh[j] = 0.f;
  if ((MODE == 2 || MODE == 3) && active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { msc[j] = scale[c + j]; msh[j] = shift[c + j]; }
  }
  // per-channel constants -> registers (two 128-bit loads per array)
  float k0[8], k1[8], k2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) k0[j] = k1[j] = k2[j] = 0.f;
  if (!active) {
  } else if (MODE == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { k0[j] = p0[c + j]; k1[j] = p1[c + j]; }             // scale, shift
  } else if (MODE == 2) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float invstd = p2[c + j], kk = p0[c + j] * invstd;                         // gamma * invstd
      // dz = kk*g - kk*mg - kk*mgx*xhat,  xhat = (z - mean) * invstd
      k0[j] = kk;
      k1[j] = kk * fsums[C + c + j] * invstd;                                          // coefficient of (z - mean)
      k2[j] = kk * fsums[c + j] - k1[j] * p1[c + j];                                   // constant part, mean folded in
    }
  } else if (MODE == 3) {
#pragma unroll
    for (int j = 0; j < 8; ++j) k0[j] = p0[c + j] * p2[c + j];
  }
  const int64_t step = static_cast<int64_t>(gridDim.x) * lanes;
  for (int64_t s = static_cast<int64_t>(blockIdx.x) * lanes + sl; active && s < n_slots; s += kUnroll * step) {
    // kUnroll slots per iteration, all loads first (halo slots are loaded too and zeroed afterwards: no divergence)
    uint4 va[kUnroll], vb[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t su = s + u * step;
      va[u] = make_uint4(0u, 0u, 0u, 0u);
      vb[u] = va[u];
      if (su < n_slots) {
        va[u] = ld8(a, su, ld_a, c);
        if (MODE == 1) vb[u] = ld8(y, su, ld_y, c);
        if (MODE == 2 || MODE == 3) vb[u] = ld8(z, su, ld_z, c);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t su = s + u * step;
      if (su >= n_slots) break;
      float r[8], av[8];
      if (MODE == 4) {
        unpack8(va[u], av, dt_yz);
        st8(out, su, ld_out, c, pack8(av, dt_a));
        continue;
      }
      const bool halo = (MODE == 0 || MODE == 2) && !slot_valid_fast(su, dv);
      unpack8(va[u], av, dt_a);
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = halo ? 0.f : fmaxf(fmaf(av[j], k0[j], k1[j]), 0.f);
        if (out2) st8(out2, su, ld_out2, c, pack8(r, dt_out2));
      } else if (MODE == 1) {
        float yv[8];
        unpack8(vb[u], yv, dt_yz);
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = yv[j] > 0.f ? av[j] : 0.f;
      } else {
        float zv[8];
        unpack8(vb[u], zv, dt_yz);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float gg = fmaf(zv[j], msc[j], msh[j]) > 0.f ? av[j] : 0.f;
          const float d = MODE == 3 ? gg * k0[j] : fmaf(k0[j], gg, -fmaf(k1[j], zv[j], k2[j]));
          r[j] = halo ? 0.f : d;
        }
      }
      const uint4 pk = pack8(r, dt_a);
      st8(out, su, ld_out, c, pk);
      if ((MODE == 2 || MODE == 3) && csum) {
        float rr[8];
        unpack8(pk, rr, dt_a);
#pragma unroll
        for (int j = 0; j < 8; ++j) cs[j] += rr[j];
      }
    }
  }
  if ((MODE == 2 || MODE == 3) && csum) {
    const int stride = groups * 8;
    if (active) {
#pragma unroll
      for (int j = 0; j < 8; ++j) map_red[sl * stride + c + j] = cs[j];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < stride; i += blockDim.x) {
      float acc = 0.f;
      for (int l = 0; l < lanes; ++l) acc += map_red[l * stride + i];
      atomicAdd(&csum[i], acc);
    }
  }
}

// per-channel means of the two backward reductions in fp32 (+ the parameter gradients dgamma = sum g*xhat, dbeta = sum g)
// (also pads gamma to the channel pitch: fsums[2 C + c] = c < C_real ? gamma[c] : 0, read by the slot pass)
__global__ void bn_bwd_means_kernel(const double* __restrict__ sums, double inv_count, int C_real, int C,
                                    float* __restrict__ fsums, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                    int accumulate, const float* __restrict__ gamma,
                                    const float* __restrict__ invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  // invstd != NULL: sums[1] holds sum g * (z - mean) (from the data-gradient conv epilogue); xhat = (z - mean) * invstd
  const double sgx = invstd ? sums[C + c] * static_cast<double>(invstd[c]) : sums[C + c];
  fsums[c] = static_cast<float>(sums[c] * inv_count);
  fsums[C + c] = static_cast<float>(sgx * inv_count);
  fsums[2 * C + c] = c < C_real ? gamma[c] : 0.0f;
  if (c < C_real && dgamma && dbeta) {
    const float db = static_cast<float>(sums[c]), dg = static_cast<float>(sgx);
    dbeta[c] = accumulate ? dbeta[c] + db : db;
    dgamma[c] = accumulate ? dgamma[c] + dg : dg;
  }
}

static SlotDiv make_div(int Hp, int Wp, int64_t n_slots) {
  SlotDiv d;
  d.per_img = static_cast<uint32_t>(Hp) * Wp;
  d.Wp = static_cast<uint32_t>(Wp);
  d.rcp_img = 1.0f / static_cast<float>(d.per_img);
  d.rcp_w = 1.0f / static_cast<float>(Wp);
  d.exact = n_slots < (1 << 24);
  return d;
}
// Grid of the slot passes: enough blocks for >= 4 slots per thread, capped at `per_sm` blocks per SM.  Measured on B200
// (64 x 96 x 96 patches): the light passes (BN apply, ReLU backward, convert: 16 resident blocks per SM) are fastest with
// many short blocks; the BatchNorm-backward apply pass (119 registers: two resident blocks per SM, 40 per-channel
// constants to load and a shared-memory column reduction + atomics per block) and the column reductions gain 8-16 %
// as persistent grids of two blocks per SM (C = 280: 0.228 -> 0.205 ms and 0.142 -> 0.131 ms).
static int map_grid(int64_t n_slots, int C, int per_sm = 16) {
  const int lanes = 256 / (C / 8);
  int64_t want = ceil_div64(n_slots, static_cast<int64_t>(lanes) * 4);     // >= 4 slots per thread
  const int64_t cap = static_cast<int64_t>(sm_count()) * per_sm;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return static_cast<int>(want);
}

static int red_grid(int64_t n_slots, int lanes) {
  int64_t want = ceil_div64(n_slots, static_cast<int64_t>(lanes) * 8);
  int64_t cap = static_cast<int64_t>(sm_count()) * 2;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return static_cast<int>(want);
}

}  // namespace mmlf

using namespace mmlf;

#define CHECK_C(C) MMLF_REQUIRE((C) % 8 == 0 && (C) >= 8 && (C) <= 2048, "channel count %d must be a multiple of 8 in [8, 2048]", (C))

// 16-bit format conversion of a slot array (fp16 activations -> bf16 operand of the weight-gradient GEMM, whose two
// operands must share one format: tcgen05.mma kind::f16 rejects mixed f16 x bf16 with an illegal-instruction fault).
extern "C" int mmlf_convert16(const void* src, int ld_src, int src_dtype, void* dst, int ld_dst, int dst_dtype, int C,
                              int64_t n_slots, void* stream) {
  MMLF_REQUIRE(src && dst, "convert16: null buffer");
  CHECK_C(C);
  slot_map_kernel<4><<<map_grid(n_slots, C), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, ld_src, nullptr, 0, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, C, make_div(1, 1, n_slots),
      n_slots, dst, ld_dst, nullptr, 0, 0, nullptr, dst_dtype, src_dtype);
  return check_launch("convert16");
}

extern "C" int mmlf_bn_stats(const void* z, int ld, int C, int B, int H, int W, int act_dtype, double* sums,
                             void* stream) {
  MMLF_REQUIRE(z && sums, "bn_stats: null buffer");
  CHECK_C(C);
  MMLF_REQUIRE(C / 8 <= kRedThreads, "bn_stats: too many channels");
  const int64_t n_slots = static_cast<int64_t>(B) * (H + 1) * (W + 1);
  const int groups = C / 8, lanes = kRedThreads / groups;
  const size_t smem = static_cast<size_t>(lanes) * groups * 16 * sizeof(float);
  col_reduce_kernel<0><<<red_grid(n_slots, lanes), kRedThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      z, ld, nullptr, nullptr, nullptr, 0, nullptr, nullptr, C, n_slots, sums, nullptr, act_dtype, 0);
  return check_launch("bn_stats");
}

extern "C" int mmlf_colsum16(const void* x, int ld, int C, int64_t n_slots, int dtype, float* out, int accumulate,
                             void* stream) {
  MMLF_REQUIRE(x && out, "colsum: null buffer");
  CHECK_C(C);
  if (!accumulate) cudaMemsetAsync(out, 0, sizeof(float) * C, static_cast<cudaStream_t>(stream));
  const int groups = C / 8, lanes = kRedThreads / groups;
  const size_t smem = static_cast<size_t>(lanes) * groups * 16 * sizeof(float);
  col_reduce_kernel<1><<<red_grid(n_slots, lanes), kRedThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      x, ld, nullptr, nullptr, nullptr, 0, nullptr, nullptr, C, n_slots, nullptr, out, dtype, 0);
  return check_launch("colsum");
}

extern "C" int mmlf_bn_finalize(const double* sums, int C_real, int C, int64_t count, const float* gamma,
                                const float* beta, float* running_mean, float* running_var,
                                int64_t* num_batches_tracked, float momentum, float eps, float* scale, float* shift,
                                float* save_mean, float* save_invstd, void* stream) {
  MMLF_REQUIRE(sums && gamma && beta && scale && shift && save_mean && save_invstd, "bn_finalize: null buffer");
  MMLF_REQUIRE(count > 0 && C_real <= C, "bn_finalize: bad sizes");
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      sums, C_real, C, count, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps, scale, shift,
      save_mean, save_invstd);
  return check_launch("bn_finalize");
}

extern "C" int mmlf_bn_fold_eval(int C_real, int C, const float* gamma, const float* beta, const float* running_mean,
                                 const float* running_var, const float* conv_bias, float eps, float* scale,
                                 float* shift, void* stream) {
  MMLF_REQUIRE(gamma && beta && running_mean && running_var && scale && shift, "bn_fold_eval: null buffer");
  bn_fold_eval_kernel<<<ceil_div(C, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      C_real, C, gamma, beta, running_mean, running_var, conv_bias, eps, scale, shift);
  return check_launch("bn_fold_eval");
}

extern "C" int mmlf_bn_apply_relu(const void* z, int ld_z, const float* scale, const float* shift, int C, int B,
                                  int H, int W, int act_dtype, void* y, int ld_y, void* y2, int ld_y2, int y2_dtype,
                                  void* stream) {
  MMLF_REQUIRE(z && scale && shift && y, "bn_apply_relu: null buffer");
  CHECK_C(C);
  MMLF_REQUIRE(!y2 || (ld_y2 >= C && ld_y2 % 8 == 0), "bn_apply_relu: bad ld_y2 %d", ld_y2);
  const int64_t n_slots = static_cast<int64_t>(B) * (H + 1) * (W + 1);
  slot_map_kernel<0><<<map_grid(n_slots, C), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      z, ld_z, nullptr, 0, nullptr, 0, scale, shift, nullptr, nullptr, nullptr, nullptr, C,
      make_div(H + 1, W + 1, n_slots), n_slots, y, ld_y, y2, ld_y2, y2_dtype, nullptr, act_dtype, act_dtype);
  return check_launch("bn_apply_relu");
}

extern "C" int mmlf_relu_bwd(const void* dy, int ld_dy, const void* y, int ld_y, int C, int64_t n_slots,
                             int grad_dtype, int act_dtype, void* dz, int ld_dz, void* stream) {
  MMLF_REQUIRE(dy && y && dz, "relu_bwd: null buffer");
  CHECK_C(C);
  slot_map_kernel<1><<<map_grid(n_slots, C), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dy, ld_dy, y, ld_y, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, C, make_div(1, 1, n_slots),
      n_slots, dz, ld_dz, nullptr, 0, 0, nullptr, grad_dtype, act_dtype);
  return check_launch("relu_bwd");
}

extern "C" int mmlf_bn_bwd_reduce(const void* dy, int ld_dy, const void* z, int ld_z, const float* scale,
                                  const float* shift, const float* save_mean, const float* save_invstd, int C, int B,
                                  int H, int W, int grad_dtype, int act_dtype, double* sums, void* stream) {
  MMLF_REQUIRE(dy && z && scale && shift && save_mean && save_invstd && sums, "bn_bwd_reduce: null buffer");
  CHECK_C(C);
  const int64_t n_slots = static_cast<int64_t>(B) * (H + 1) * (W + 1);
  const int groups = C / 8, lanes = kRedThreads / groups;
  const size_t smem = static_cast<size_t>(lanes) * groups * 16 * sizeof(float);
  col_reduce_kernel<2><<<red_grid(n_slots, lanes), kRedThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      dy, ld_dy, scale, shift, z, ld_z, save_mean, save_invstd, C, n_slots, sums, nullptr, grad_dtype, act_dtype);
  return check_launch("bn_bwd_reduce");
}

extern "C" int mmlf_bn_bwd_apply(const void* dy, int ld_dy, const void* z, int ld_z, const float* scale,
                                 const float* shift, const float* gamma, const float* save_mean,
                                 const float* save_invstd, const double* sums, int64_t count, int train, int C_real,
                                 int C, int B, int H, int W, int grad_dtype, int act_dtype, void* dz, int ld_dz,
                                 float* dgamma, float* dbeta, int accumulate, float* fsums, float* dz_colsum,
                                 void* stream) {
  MMLF_REQUIRE(dy && z && scale && shift && gamma && save_mean && save_invstd && sums && dz, "bn_bwd_apply: null buffer");
  CHECK_C(C);
  const int64_t n_slots = static_cast<int64_t>(B) * (H + 1) * (W + 1);
  const int blocks = map_grid(n_slots, C, 2);
  const SlotDiv dv = make_div(H + 1, W + 1, n_slots);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MMLF_REQUIRE(fsums != nullptr, "bn_bwd_apply: fsums scratch (float[3*C]) required");
  bn_bwd_means_kernel<<<ceil_div(C, 128), 128, 0, st>>>(sums, 1.0 / static_cast<double>(count), C_real, C, fsums, dgamma,
                                                         dbeta, accumulate, gamma, train == 2 ? save_invstd : nullptr);
  if (int rc = check_launch("bn_bwd_means")) return rc;
  const float* gamma_pad = fsums + 2 * C;            // gamma on the padded channel pitch, written by the kernel above
  const size_t smem = dz_colsum ? sizeof(float) * 256 * 8 : 0;
  if (train)
    slot_map_kernel<2><<<blocks, 256, smem, st>>>(dy, ld_dy, nullptr, 0, z, ld_z, gamma_pad, save_mean, save_invstd, fsums,
                                                  scale, shift, C, dv, n_slots, dz, ld_dz, nullptr, 0, 0, dz_colsum,
                                                  grad_dtype, act_dtype);
  else
    slot_map_kernel<3><<<blocks, 256, smem, st>>>(dy, ld_dy, nullptr, 0, z, ld_z, gamma_pad, save_mean, save_invstd, fsums,
                                                  scale, shift, C, dv, n_slots, dz, ld_dz, nullptr, 0, 0, dz_colsum,
                                                  grad_dtype, act_dtype);
  return check_launch("bn_bwd_apply");
}
