// Generic float32 layer kernels for the topologies outside the published 2x2 / 8-block network:
// odd --model_ksize (3x3 convolutions with symmetric padding, /root/reference/mmlf/model/feed_forward.py:86-92) and the
// --model_unet out-net (/root/reference/mmlf/model/unet.py:8-132: 3x3 conv -> ReLU -> BatchNorm blocks, 2x2 max-pool,
// stride-2 transposed convolution, centre crop + concat, 1x1 head).
//
// Layout: dense channel-last float32, a tensor of B x H x W pixels with C channels is [B*H*W][ld] (ld >= C, so that a
// tensor can be a channel slice of a wider one).  Everything is CUDA-core fp32 (fp32 accumulate): these rows are about
// coverage and exact parity, the tensor-core path is the published topology's (conv2x2_tc.cu).
//   convolution  : implicit GEMM, 64 x 64 x 16 shared-memory tiles, 4 x 4 register micro-tiles; the data gradient is the
//                  same kernel on rotated / transposed weights (mmlf_g_pack_weight) with padding k - 1 - pad
//   weight grad  : [k*k*cin][cout] tiles over a slice of the pixels, atomically added into the canonical (cout, cin, k, k)
//                  gradient (which also accumulates the two calls of a shared in-net)
//   transposed conv (k 2, stride 2) = 1x1 convolution to 4 * cout channels + depth-to-space
#include "../../include/mmlf_b200.h"
#include "common.cuh"
#include "host_util.h"

namespace mmlf {

constexpr int GBM = 64, GBN = 64, GBK = 16;

// canonical tap (u, v) of the effective tap (dy, dx) for the stream plumbing folded into the weights
// (feed_forward.py:236-256): 0 none, 1 transpose (h stream), 2 transpose + flip (i stream): w'[dy][dx] = w[dx][k-1-dy]
__device__ __forceinline__ void canon_tap(int spatial, int k, int dy, int dx, int& u, int& v) {
  if (spatial == 0) { u = dy; v = dx; }
  else if (spatial == 1) { u = dx; v = dy; }
  else { u = dx; v = k - 1 - dy; }
}

// w: canonical (cout, cin, k, k) [or (cin, cout, 2, 2) for transposed = 2].  out: GEMM operand [K][N]:
//   mode 0 forward : K = (dy*k+dx)*cin + ci, N = cout
//   mode 1 dgrad   : K = ((k-1-dy)*k + (k-1-dx))*cout + co, N = cin            (effective taps rotated by 180 degrees)
//   mode 2 convT fwd: w (cin, cout, 2, 2): K = ci, N = (dy*2+dx)*cout + co
//   mode 3 convT dgrad: K = (dy*2+dx)*cout + co, N = ci
__global__ void g_pack_weight_kernel(const float* __restrict__ w, int cout, int cin, int k, int spatial, int mode,
                                     float* __restrict__ out) {
  const int total = cout * cin * k * k;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    if (mode <= 1) {
      // enumerate effective taps
      const int ci = i % cin, co = (i / cin) % cout, tap = i / (cin * cout);
      const int dy = tap / k, dx = tap % k;
      int u, v;
      canon_tap(spatial, k, dy, dx, u, v);
      const float val = w[((static_cast<int64_t>(co) * cin + ci) * k + u) * k + v];
      if (mode == 0) out[(static_cast<int64_t>(tap) * cin + ci) * cout + co] = val;
      else out[(static_cast<int64_t>((k - 1 - dy) * k + (k - 1 - dx)) * cout + co) * cin + ci] = val;
    } else {
      const int co = i % cout, ci = (i / cout) % cin, tap = i / (cin * cout);       // w[ci][co][dy][dx], k = 2
      const float val = w[(static_cast<int64_t>(ci) * cout + co) * 4 + tap];
      if (mode == 2) out[static_cast<int64_t>(ci) * (4 * cout) + tap * cout + co] = val;
      else out[(static_cast<int64_t>(tap) * cout + co) * cin + ci] = val;
    }
  }
}

// y[m][n] = act( sum_kk A[m][kk] * w[kk][n] + bias[n] ), A gathered from x with zero padding.
__global__ void __launch_bounds__(256)
g_conv_kernel(const float* __restrict__ x, int ld_x, const float* __restrict__ w, const float* __restrict__ bias, int B,
              int H, int W, int cin, int Ho, int Wo, int N, int k, int pad, int relu, float* __restrict__ y, int ld_y) {
  __shared__ float As[GBK][GBM + 4];
  __shared__ float Bs[GBK][GBN + 4];
  const int64_t M = static_cast<int64_t>(B) * Ho * Wo;
  const int K = k * k * cin;
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * GBM;
  const int n0 = blockIdx.y * GBN;
  const int t = threadIdx.x;
  const int a_kk = t % GBK, a_m = t / GBK;                 // A loads: 4 rows per thread, kk fastest (coalesced over ci)
  int rb[4], ry[4], rx[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + a_m + 16 * i;
    if (m < M) {
      const int64_t img = m / (static_cast<int64_t>(Ho) * Wo);
      const int rem = static_cast<int>(m - img * Ho * Wo);
      rb[i] = static_cast<int>(img); ry[i] = rem / Wo; rx[i] = rem % Wo;
    } else {
      rb[i] = -1; ry[i] = 0; rx[i] = 0;
    }
  }
  const int b_n = t % GBN, b_kk = t / GBN;                 // B loads: 4 k-rows per thread, n fastest
  const int tm = (t / 16) * 4, tn = (t % 16) * 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += GBK) {
    const int kk = k0 + a_kk;
    int dy = 0, dx = 0, ci = 0;
    const bool kin = kk < K;
    if (kin) {
      const int tap = kk / cin;
      ci = kk - tap * cin;
      dy = tap / k; dx = tap - dy * k;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float v = 0.f;
      if (kin && rb[i] >= 0) {
        const int iy = ry[i] + dy - pad, ix = rx[i] + dx - pad;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W)
          v = __ldg(x + ((static_cast<int64_t>(rb[i]) * H + iy) * W + ix) * ld_x + ci);
      }
      As[a_kk][a_m + 16 * i] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kr = k0 + b_kk + 4 * i, n = n0 + b_n;
      Bs[b_kk + 4 * i][b_n] = (kr < K && n < N) ? __ldg(w + static_cast<int64_t>(kr) * N + n) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < GBK; ++q) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[q][tm + i]; b[i] = Bs[q][tn + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + tm + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tn + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (relu) v = fmaxf(v, 0.f);
      y[m * ld_y + n] = v;
    }
  }
}

// dw[co][ci][u][v] += sum over the block's pixel slice of A[m][(tap, ci)] * dy[m][co]   (atomics; dw zeroed by the caller)
__global__ void __launch_bounds__(256)
g_wgrad_kernel(const float* __restrict__ x, int ld_x, const float* __restrict__ dy, int ld_dy, int B, int H, int W,
               int cin, int Ho, int Wo, int cout, int k, int pad, int spatial, int transposed, int64_t m_per_block,
               float* __restrict__ dw) {
  __shared__ float As[GBK][GBM + 4];                      // [m][kk]
  __shared__ float Ds[GBK][GBN + 4];                      // [m][n]
  const int64_t M = static_cast<int64_t>(B) * Ho * Wo;
  const int K = k * k * cin;
  const int kb = blockIdx.x * GBM, n0 = blockIdx.y * GBN;
  const int64_t mlo = static_cast<int64_t>(blockIdx.z) * m_per_block;
  const int64_t mhi = mlo + m_per_block < M ? mlo + m_per_block : M;
  const int t = threadIdx.x;
  const int l_c = t % 64, l_m = t / 64;                   // loads: 4 pixel rows per thread, column fastest
  const int kk = kb + l_c;
  int tdy = 0, tdx = 0, ci = 0;
  const bool kin = kk < K;
  if (kin) {
    const int tap = kk / cin;
    ci = kk - tap * cin;
    tdy = tap / k; tdx = tap - tdy * k;
  }
  const int tk = (t / 16) * 4, tn = (t % 16) * 4;
  float acc[4][4] = {};
  for (int64_t mb = mlo; mb < mhi; mb += GBK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t m = mb + l_m + 4 * i;
      float a = 0.f, d = 0.f;
      if (m < mhi) {
        const int64_t img = m / (static_cast<int64_t>(Ho) * Wo);
        const int rem = static_cast<int>(m - img * Ho * Wo);
        const int oy = rem / Wo, ox = rem % Wo;
        if (kin) {
          const int iy = oy + tdy - pad, ix = ox + tdx - pad;
          if (iy >= 0 && iy < H && ix >= 0 && ix < W) a = __ldg(x + ((img * H + iy) * W + ix) * ld_x + ci);
        }
        if (n0 + l_c < cout) d = __ldg(dy + m * ld_dy + n0 + l_c);
      }
      As[l_m + 4 * i][l_c] = a;
      Ds[l_m + 4 * i][l_c] = d;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < GBK; ++q) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[q][tk + i]; b[i] = Ds[q][tn + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int kq = kb + tk + i;
    if (kq >= K) continue;
    const int tap = kq / cin, c = kq - tap * cin;
    const int dyy = tap / k, dxx = tap - dyy * k;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tn + j;
      if (n >= cout) continue;
      int64_t off;
      if (transposed) {                                    // x: (.., cin), dy: (.., 4 * cout_t): dw (cin, cout_t, 2, 2)
        const int cout_t = cout / 4, tp = n / cout_t, co = n - tp * cout_t;
        off = (static_cast<int64_t>(c) * cout_t + co) * 4 + tp;
      } else {
        int u, v;
        canon_tap(spatial, k, dyy, dxx, u, v);
        off = ((static_cast<int64_t>(n) * cin + c) * k + u) * k + v;
      }
      atomicAdd(dw + off, acc[i][j]);
    }
  }
}

// out[c] += sum_rows x[row][c]   (bias gradients; `fold`: channels c, c + C, ... of a 4*C-wide row add into out[c])
__global__ void g_colsum_kernel(const float* __restrict__ x, int ld, int C, int64_t n_rows, float* __restrict__ out) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int lane_row = threadIdx.x >> 5;                  // 8 row lanes per block
  __shared__ float red[8][33];
  float s = 0.f;
  if (c < C)
    for (int64_t r = static_cast<int64_t>(blockIdx.y) * 8 + lane_row; r < n_rows; r += static_cast<int64_t>(gridDim.y) * 8)
      s += x[r * ld + c];
  red[lane_row][threadIdx.x & 31] = s;
  __syncthreads();
  if (lane_row == 0 && c < C) {
    float tsum = 0.f;
    for (int i = 0; i < 8; ++i) tsum += red[i][threadIdx.x & 31];
    atomicAdd(out + c, tsum);
  }
}

// BatchNorm statistics: sums[c] += sum x, sums[C + c] += sum x^2 (double)
__global__ void g_bn_stats_kernel(const float* __restrict__ x, int ld, int C, int64_t n_rows, double* __restrict__ sums) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int lane_row = threadIdx.x >> 5;
  __shared__ double r1[8][33], r2[8][33];
  double s = 0.0, s2 = 0.0;
  if (c < C)
    for (int64_t r = static_cast<int64_t>(blockIdx.y) * 8 + lane_row; r < n_rows; r += static_cast<int64_t>(gridDim.y) * 8) {
      const double v = static_cast<double>(x[r * ld + c]);
      s += v; s2 += v * v;
    }
  r1[lane_row][threadIdx.x & 31] = s;
  r2[lane_row][threadIdx.x & 31] = s2;
  __syncthreads();
  if (lane_row == 0 && c < C) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < 8; ++i) { a += r1[i][threadIdx.x & 31]; b += r2[i][threadIdx.x & 31]; }
    atomicAdd(sums + c, a);
    atomicAdd(sums + C + c, b);
  }
}

// y = x * scale + shift [relu]
__global__ void g_affine_kernel(const float* __restrict__ x, int ld_x, const float* __restrict__ scale,
                                const float* __restrict__ shift, int C, int64_t n_rows, int relu, float* __restrict__ y,
                                int ld_y) {
  const int64_t total = n_rows * C;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / C;
    const int c = static_cast<int>(i - r * C);
    float v = fmaf(x[r * ld_x + c], scale[c], shift[c]);
    if (relu) v = fmaxf(v, 0.f);
    y[r * ld_y + c] = v;
  }
}

// BatchNorm backward statistics: with g = dy * (gate ? gate > 0 : 1) and xhat = (x - mean) * invstd:
// sums[c] += sum g, sums[C + c] += sum g * xhat
__global__ void g_bn_bwd_reduce_kernel(const float* __restrict__ dy, int ld_dy, const float* __restrict__ x, int ld_x,
                                       const float* __restrict__ gate, int ld_gate, const float* __restrict__ mean,
                                       const float* __restrict__ invstd, int C, int64_t n_rows,
                                       double* __restrict__ sums) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int lane_row = threadIdx.x >> 5;
  __shared__ double r1[8][33], r2[8][33];
  double s = 0.0, s2 = 0.0;
  if (c < C) {
    const float mu = mean[c], is = invstd[c];
    for (int64_t r = static_cast<int64_t>(blockIdx.y) * 8 + lane_row; r < n_rows; r += static_cast<int64_t>(gridDim.y) * 8) {
      float g = dy[r * ld_dy + c];
      if (gate && !(gate[r * ld_gate + c] > 0.f)) g = 0.f;
      s += static_cast<double>(g);
      s2 += static_cast<double>(g * ((x[r * ld_x + c] - mu) * is));
    }
  }
  r1[lane_row][threadIdx.x & 31] = s;
  r2[lane_row][threadIdx.x & 31] = s2;
  __syncthreads();
  if (lane_row == 0 && c < C) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < 8; ++i) { a += r1[i][threadIdx.x & 31]; b += r2[i][threadIdx.x & 31]; }
    atomicAdd(sums + c, a);
    atomicAdd(sums + C + c, b);
  }
}

// dx = gamma * invstd * (g - sum_g / n - xhat * sum_gx / n)   (train)   |   g * gamma * invstd   (eval-mode BatchNorm)
__global__ void g_bn_bwd_apply_kernel(const float* __restrict__ dy, int ld_dy, const float* __restrict__ x, int ld_x,
                                      const float* __restrict__ gate, int ld_gate, const float* __restrict__ gamma,
                                      const float* __restrict__ mean, const float* __restrict__ invstd,
                                      const double* __restrict__ sums, double inv_count, int train, int C,
                                      int64_t n_rows, float* __restrict__ dx, int ld_dx) {
  const int64_t total = n_rows * C;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / C;
    const int c = static_cast<int>(i - r * C);
    float g = dy[r * ld_dy + c];
    if (gate && !(gate[r * ld_gate + c] > 0.f)) g = 0.f;
    const float kk = gamma[c] * invstd[c];
    float v;
    if (train) {
      const float xhat = (x[r * ld_x + c] - mean[c]) * invstd[c];
      v = kk * (g - static_cast<float>(sums[c] * inv_count) - xhat * static_cast<float>(sums[C + c] * inv_count));
    } else {
      v = g * kk;
    }
    dx[r * ld_dx + c] = v;
  }
}

// dgamma (+)= sums[C + c], dbeta (+)= sums[c]
__global__ void g_bn_param_grads_kernel(const double* __restrict__ sums, int C, float* __restrict__ dgamma,
                                        float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  dbeta[c] += static_cast<float>(sums[c]);
  dgamma[c] += static_cast<float>(sums[C + c]);
}

// out = dy * (y > 0)
__global__ void g_relu_bwd_kernel(const float* __restrict__ dy, int ld_dy, const float* __restrict__ y, int ld_y, int C,
                                  int64_t n_rows, float* __restrict__ out, int ld_out) {
  const int64_t total = n_rows * C;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / C;
    const int c = static_cast<int>(i - r * C);
    out[r * ld_out + c] = y[r * ld_y + c] > 0.f ? dy[r * ld_dy + c] : 0.f;
  }
}

// F.max_pool2d(x, 2): (B, H, W, C) -> (B, H/2, W/2, C); idx = position of the first maximum in the row-major window
__global__ void g_maxpool_kernel(const float* __restrict__ x, int B, int H, int W, int C, float* __restrict__ y,
                                 uint8_t* __restrict__ idx) {
  const int h = H / 2, w = W / 2;
  const int64_t total = static_cast<int64_t>(B) * h * w * C;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const int64_t p = i / C;
    const int ox = static_cast<int>(p % w), oy = static_cast<int>((p / w) % h);
    const int64_t b = p / (static_cast<int64_t>(w) * h);
    float best = 0.f;
    int bi = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float v = x[((b * H + 2 * oy + (q >> 1)) * W + 2 * ox + (q & 1)) * C + c];
      if (q == 0 || v > best || (v != v && best == best)) { best = v; bi = q; }     // NaN propagates like PyTorch
    }
    y[i] = best;
    idx[i] = static_cast<uint8_t>(bi);
  }
}

// dx (B, H, W, C), every element written: the pooled gradient at the arg-max position, zero elsewhere (odd borders too)
__global__ void g_maxpool_bwd_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ idx, int B, int H, int W,
                                     int C, float* __restrict__ dx) {
  const int h = H / 2, w = W / 2;
  const int64_t total = static_cast<int64_t>(B) * H * W * C;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const int64_t p = i / C;
    const int xx = static_cast<int>(p % W), yy = static_cast<int>((p / W) % H);
    const int64_t b = p / (static_cast<int64_t>(W) * H);
    const int oy = yy / 2, ox = xx / 2;
    float v = 0.f;
    if (oy < h && ox < w) {
      const int64_t o = ((b * h + oy) * w + ox) * C + c;
      if (idx[o] == ((yy & 1) << 1 | (xx & 1))) v = dy[o];
    }
    dx[i] = v;
  }
}

// window copy between channel-last tensors: dst[b][yd0+y][xd0+x][cd0+c] (+)= src[b][ys0+y][xs0+x][cs0+c]
__global__ void g_copy_window_kernel(const float* __restrict__ src, int Hs, int Ws, int ld_s, int cs0, int ys0, int xs0,
                                     float* __restrict__ dst, int Hd, int Wd, int ld_d, int cd0, int yd0, int xd0, int B,
                                     int h, int w, int C, int accumulate) {
  const int64_t total = static_cast<int64_t>(B) * h * w * C;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const int64_t p = i / C;
    const int x = static_cast<int>(p % w), y = static_cast<int>((p / w) % h);
    const int64_t b = p / (static_cast<int64_t>(w) * h);
    const float v = src[((b * Hs + ys0 + y) * Ws + xs0 + x) * ld_s + cs0 + c];
    float* d = dst + ((b * Hd + yd0 + y) * Wd + xd0 + x) * ld_d + cd0 + c;
    *d = accumulate ? *d + v : v;
  }
}

// depth-to-space (dir 0): y4 (B, H, W, 4*C) [tap = dy*2+dx major] -> out (B, 2H, 2W, C) at channel offset c0 of pitch ld;
// space-to-depth (dir 1): the inverse gather.
__global__ void g_d2s_kernel(float* __restrict__ y4, float* __restrict__ out, int ld, int c0, int B, int H, int W, int C,
                             int dir) {
  const int64_t total = static_cast<int64_t>(B) * H * W * 4 * C;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const int tap = static_cast<int>((i / C) % 4);
    const int64_t p = i / (4 * C);
    const int x = static_cast<int>(p % W), y = static_cast<int>((p / W) % H);
    const int64_t b = p / (static_cast<int64_t>(W) * H);
    float* o = out + ((b * 2 * H + 2 * y + (tap >> 1)) * 2 * W + 2 * x + (tap & 1)) * ld + c0 + c;
    if (dir == 0) *o = y4[i];
    else y4[i] = *o;
  }
}

// (B, C, H, W) <-> (B, H, W, C at pitch ld); `spatial` is not applied here (it lives in the weights)
__global__ void g_layout_kernel(float* __restrict__ nchw, float* __restrict__ nhwc, int ld, int B, int C, int H, int W,
                                int to_nhwc) {
  const int64_t total = static_cast<int64_t>(B) * C * H * W;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const int64_t p = i / C;                                                  // pixel index (b, y, x)
    const int64_t b = p / (static_cast<int64_t>(H) * W), yx = p - b * H * W;
    float* a = nchw + (b * C + c) * H * W + yx;
    float* d = nhwc + p * ld + c;
    if (to_nhwc) *d = *a;
    else *a = *d;
  }
}

static int ew_grid(int64_t total) {
  int64_t g = ceil_div64(total, 256 * 4);
  const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
  if (g > cap) g = cap;
  return g < 1 ? 1 : static_cast<int>(g);
}

static dim3 col_grid(int C, int64_t n_rows) {
  int64_t gy = ceil_div64(n_rows, 8 * 32);
  const int64_t cap = (static_cast<int64_t>(sm_count()) * 8) / ceil_div(C, 32) + 1;
  if (gy > cap) gy = cap;
  if (gy < 1) gy = 1;
  return dim3(ceil_div(C, 32), static_cast<unsigned>(gy));
}

}  // namespace mmlf

using namespace mmlf;

extern "C" int mmlf_g_pack_weight(const float* w, int cout, int cin, int k, int spatial, int mode, float* out,
                                  void* stream) {
  MMLF_REQUIRE(w && out && k >= 1 && k <= 7 && mode >= 0 && mode <= 3 && spatial >= 0 && spatial <= 2, "g_pack_weight: bad arguments");
  MMLF_REQUIRE(mode <= 1 || k == 2, "g_pack_weight: transposed convolutions are k = 2, stride 2");
  g_pack_weight_kernel<<<ew_grid(static_cast<int64_t>(cout) * cin * k * k), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, cout, cin, k, spatial, mode, out);
  return check_launch("g_pack_weight");
}

extern "C" int mmlf_g_conv(const float* x, int ld_x, const float* wg, const float* bias, int B, int H, int W, int cin,
                           int cout, int k, int pad, int relu, float* y, int ld_y, void* stream) {
  MMLF_REQUIRE(x && wg && y, "g_conv: null buffer");
  const int Ho = H + 2 * pad - k + 1, Wo = W + 2 * pad - k + 1;
  MMLF_REQUIRE(Ho >= 1 && Wo >= 1 && ld_x >= cin && ld_y >= cout, "g_conv: bad geometry");
  const int64_t M = static_cast<int64_t>(B) * Ho * Wo;
  dim3 grid(static_cast<unsigned>(ceil_div64(M, GBM)), ceil_div(cout, GBN));
  g_conv_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ld_x, wg, bias, B, H, W, cin, Ho, Wo, cout, k, pad,
                                                                     relu, y, ld_y);
  return check_launch("g_conv");
}

extern "C" int mmlf_g_conv_wgrad(const float* x, int ld_x, const float* dy, int ld_dy, int B, int H, int W, int cin,
                                 int cout, int k, int pad, int spatial, int transposed, float* dw, void* stream) {
  MMLF_REQUIRE(x && dy && dw, "g_conv_wgrad: null buffer");
  const int Ho = H + 2 * pad - k + 1, Wo = W + 2 * pad - k + 1;
  MMLF_REQUIRE(Ho >= 1 && Wo >= 1, "g_conv_wgrad: bad geometry");
  MMLF_REQUIRE(!transposed || (k == 1 && cout % 4 == 0), "g_conv_wgrad: transposed form is a 1x1 GEMM to 4 * cout channels");
  const int64_t M = static_cast<int64_t>(B) * Ho * Wo;
  const int K = k * k * cin;
  const int tiles = ceil_div(K, GBM) * ceil_div(cout, GBN);
  int64_t split = (static_cast<int64_t>(sm_count()) * 4 + tiles - 1) / tiles;
  const int64_t max_split = ceil_div64(M, 256);
  if (split > max_split) split = max_split;
  if (split < 1) split = 1;
  int64_t per = ceil_div64(M, split);
  per = ceil_div64(per, GBK) * GBK;
  split = ceil_div64(M, per);
  dim3 grid(ceil_div(K, GBM), ceil_div(cout, GBN), static_cast<unsigned>(split));
  g_wgrad_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ld_x, dy, ld_dy, B, H, W, cin, Ho, Wo, cout, k, pad,
                                                                      spatial, transposed, per, dw);
  return check_launch("g_conv_wgrad");
}

extern "C" int mmlf_g_colsum(const float* x, int ld, int C, int64_t n_rows, float* out, void* stream) {
  MMLF_REQUIRE(x && out && C >= 1, "g_colsum: bad arguments");
  g_colsum_kernel<<<col_grid(C, n_rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ld, C, n_rows, out);
  return check_launch("g_colsum");
}

extern "C" int mmlf_g_bn_stats(const float* x, int ld, int C, int64_t n_rows, double* sums, void* stream) {
  MMLF_REQUIRE(x && sums && C >= 1, "g_bn_stats: bad arguments");
  g_bn_stats_kernel<<<col_grid(C, n_rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ld, C, n_rows, sums);
  return check_launch("g_bn_stats");
}

extern "C" int mmlf_g_affine(const float* x, int ld_x, const float* scale, const float* shift, int C, int64_t n_rows,
                             int relu, float* y, int ld_y, void* stream) {
  MMLF_REQUIRE(x && scale && shift && y, "g_affine: null buffer");
  g_affine_kernel<<<ew_grid(n_rows * C), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ld_x, scale, shift, C, n_rows, relu,
                                                                                    y, ld_y);
  return check_launch("g_affine");
}

extern "C" int mmlf_g_bn_bwd(const float* dy, int ld_dy, const float* x, int ld_x, const float* gate, int ld_gate,
                             const float* gamma, const float* mean, const float* invstd, double* sums, int64_t count,
                             int train, int C, int64_t n_rows, float* dx, int ld_dx, float* dgamma, float* dbeta,
                             void* stream) {
  MMLF_REQUIRE(dy && x && gamma && mean && invstd && sums && dx && dgamma && dbeta, "g_bn_bwd: null buffer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  g_bn_bwd_reduce_kernel<<<col_grid(C, n_rows), 256, 0, st>>>(dy, ld_dy, x, ld_x, gate, ld_gate, mean, invstd, C, n_rows,
                                                             sums);
  if (int rc = check_launch("g_bn_bwd_reduce")) return rc;
  g_bn_param_grads_kernel<<<ceil_div(C, 128), 128, 0, st>>>(sums, C, dgamma, dbeta);
  if (int rc = check_launch("g_bn_param_grads")) return rc;
  g_bn_bwd_apply_kernel<<<ew_grid(n_rows * C), 256, 0, st>>>(dy, ld_dy, x, ld_x, gate, ld_gate, gamma, mean, invstd, sums,
                                                            1.0 / static_cast<double>(count), train, C, n_rows, dx, ld_dx);
  return check_launch("g_bn_bwd_apply");
}

extern "C" int mmlf_g_relu_bwd(const float* dy, int ld_dy, const float* y, int ld_y, int C, int64_t n_rows, float* out,
                               int ld_out, void* stream) {
  MMLF_REQUIRE(dy && y && out, "g_relu_bwd: null buffer");
  g_relu_bwd_kernel<<<ew_grid(n_rows * C), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, ld_dy, y, ld_y, C, n_rows, out,
                                                                                      ld_out);
  return check_launch("g_relu_bwd");
}

extern "C" int mmlf_g_maxpool2(const float* x, int B, int H, int W, int C, float* y, uint8_t* idx, void* stream) {
  MMLF_REQUIRE(x && y && idx && H >= 2 && W >= 2, "g_maxpool2: bad arguments");
  g_maxpool_kernel<<<ew_grid(static_cast<int64_t>(B) * (H / 2) * (W / 2) * C), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, B, H, W, C, y, idx);
  return check_launch("g_maxpool2");
}

extern "C" int mmlf_g_maxpool2_bwd(const float* dy, const uint8_t* idx, int B, int H, int W, int C, float* dx,
                                   void* stream) {
  MMLF_REQUIRE(dy && idx && dx, "g_maxpool2_bwd: null buffer");
  g_maxpool_bwd_kernel<<<ew_grid(static_cast<int64_t>(B) * H * W * C), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dy, idx, B, H, W, C, dx);
  return check_launch("g_maxpool2_bwd");
}

extern "C" int mmlf_g_copy_window(const float* src, int Hs, int Ws, int ld_s, int cs0, int ys0, int xs0, float* dst,
                                  int Hd, int Wd, int ld_d, int cd0, int yd0, int xd0, int B, int h, int w, int C,
                                  int accumulate, void* stream) {
  MMLF_REQUIRE(src && dst, "g_copy_window: null buffer");
  MMLF_REQUIRE(ys0 >= 0 && xs0 >= 0 && ys0 + h <= Hs && xs0 + w <= Ws && yd0 >= 0 && xd0 >= 0 && yd0 + h <= Hd &&
                   xd0 + w <= Wd && cs0 + C <= ld_s && cd0 + C <= ld_d,
               "g_copy_window: window outside the tensors");
  g_copy_window_kernel<<<ew_grid(static_cast<int64_t>(B) * h * w * C), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, Hs, Ws, ld_s, cs0, ys0, xs0, dst, Hd, Wd, ld_d, cd0, yd0, xd0, B, h, w, C, accumulate);
  return check_launch("g_copy_window");
}

extern "C" int mmlf_g_depth_to_space(float* y4, float* out, int ld, int c0, int B, int H, int W, int C, int inverse,
                                     void* stream) {
  MMLF_REQUIRE(y4 && out && c0 + C <= ld, "g_depth_to_space: bad arguments");
  g_d2s_kernel<<<ew_grid(static_cast<int64_t>(B) * H * W * 4 * C), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      y4, out, ld, c0, B, H, W, C, inverse);
  return check_launch("g_depth_to_space");
}

extern "C" int mmlf_g_layout(float* nchw, float* nhwc, int ld, int B, int C, int H, int W, int to_nhwc, void* stream) {
  MMLF_REQUIRE(nchw && nhwc && ld >= C, "g_layout: bad arguments");
  g_layout_kernel<<<ew_grid(static_cast<int64_t>(B) * C * H * W), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      nchw, nhwc, ld, B, C, H, W, to_nhwc);
  return check_launch("g_layout");
}
