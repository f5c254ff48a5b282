// Light-field input kernels: view-index extraction (u8 -> f32), the disparity Shift resampler and the
// conversions into the bf16 slot layout.  Bandwidth-bound; bit-exact with the reference
// (/root/reference/mmlf/data/hci4d.py:142-193, 907-990).
#include <math.h>
#include <stdlib.h>

#include "../../include/mmlf_b200.h"
#include "common.cuh"
#include "host_util.h"

namespace mmlf {

// ------------------------------------------------------------------------------------------------ extract
struct ExtractParams {
  int idx[4][16];   // view index per stack / position (hci4d.py:142-149)
  float* dst[4];
  float* center;
  int n, H, W;
};

// Correctly rounded x / 255 for a byte value x without the IEEE-division slow path: q = x * fl(1 / 255), one Newton
// correction through two FMAs.  Equal to __fdiv_rn(x, 255) for all 256 inputs (checked exhaustively in exact rational
// arithmetic, and by tests/test_gpu_kernels.py::test_lf_extract_bit_exact which feeds every byte value).
__device__ __forceinline__ float div255(uint8_t b) {
  const float x = static_cast<float>(b), rc = 0.003921568859368563f;
  const float q = __fmul_rn(x, rc);
  const float r = __fmaf_rn(-q, 255.f, x);
  return __fmaf_rn(r, rc, q);
}

__global__ void lf_extract_kernel(const uint8_t* __restrict__ views, const ExtractParams p) {
  // one thread = 4 consecutive pixels of one view of one stack
  const int W4 = p.W >> 2;
  const int64_t per_view = static_cast<int64_t>(p.H) * W4;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int sv = blockIdx.y;                     // stack * n + view position
  if (idx >= per_view) return;
  const int stack = sv / p.n, k = sv - stack * p.n;
  const int y = static_cast<int>(idx / W4), x = static_cast<int>(idx - static_cast<int64_t>(y) * W4) * 4;
  const int64_t plane = static_cast<int64_t>(p.H) * p.W;
  const uint8_t* src = views + (static_cast<int64_t>(p.idx[stack][k]) * plane + static_cast<int64_t>(y) * p.W + x) * 3;
  const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);   // 12 bytes, 4-byte aligned (x % 4 == 0)
  const uint32_t w0 = __ldg(s32), w1 = __ldg(s32 + 1), w2 = __ldg(s32 + 2);
  uint8_t px[12];
  *reinterpret_cast<uint32_t*>(px) = w0;
  *reinterpret_cast<uint32_t*>(px + 4) = w1;
  *reinterpret_cast<uint32_t*>(px + 8) = w2;
  float* dst = p.dst[stack] + static_cast<int64_t>(k) * 3 * plane + static_cast<int64_t>(y) * p.W + x;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float4 v;
    // img_as_float(u8).astype(f32) == correctly rounded x / 255 (hci4d.py:156-157)
    v.x = div255(px[c]);
    v.y = div255(px[3 + c]);
    v.z = div255(px[6 + c]);
    v.w = div255(px[9 + c]);
    __stcs(reinterpret_cast<float4*>(dst + c * plane), v);
    if (p.center && stack == 1 && k == p.n / 2)
      *reinterpret_cast<float4*>(p.center + c * plane + static_cast<int64_t>(y) * p.W + x) = v;
  }
}

// ------------------------------------------------------------------------------------------------ shift
struct ShiftTaps {
  float w0[16], w1[16];
  int s0[16], s1[16];
};

static void host_taps(double disp, int n, ShiftTaps& t) {
  const int c = n / 2;
  for (int i = 0; i < n; ++i) {
    // alpha, shift0 = math.modf(disp * (i - c)); alpha = |alpha|; shift1 = shift0 + copysign(1, shift0)
    double ip;
    double alpha = modf(disp * static_cast<double>(i - c), &ip);
    alpha = fabs(alpha);
    const double s1 = ip + copysign(1.0, ip);
    t.w0[i] = static_cast<float>(1.0 - alpha);
    t.w1[i] = static_cast<float>(alpha);
    t.s0[i] = static_cast<int>(ip);
    t.s1[i] = static_cast<int>(s1);
  }
}

// Source index of cat([x[-s:], x[:-s]]) (sign=+1) / cat([x[s:], x[:s]]) (sign=-1): circular for 0 < |s| < n,
// identity otherwise (python slice clamping).
__device__ __forceinline__ int src_index(int j, int s, int n, int sign) {
  if (s == 0 || s >= n || -s >= n) return j;
  int r = j - sign * s;
  if (r < 0) r += n;
  if (r >= n) r -= n;
  return r;
}

struct ShiftParams {
  const float* src[4];
  float* dst[4];
  ShiftTaps taps;
  int batch, n, H, W;
};

// a*w0 + b*w1 with two roundings and an add, no FMA contraction (hci4d.py:940-945)
__device__ __forceinline__ float lerp2(float a, float w0, float b, float w1) {
  return __fadd_rn(__fmul_rn(a, w0), __fmul_rn(b, w1));
}

// One CTA = a band of `rows_per_cta` output rows of one (stack, batch, view, colour) plane, full width.  The source
// rows the band needs -- a circularly contiguous run of rows_per_cta (+1 when the stack is resampled along H: the two
// taps are neighbouring rows) -- are staged in shared memory with 128-bit loads, 4 in flight per thread; the shifted,
// wrapping taps are read from there and the band is written with 128-bit stores.  Bands are sized to move >= 64 KB
// per CTA: the first version (one CTA per output row) moved 0.4 - 2 KB per CTA and reached 16 % (96-px patches) /
// 48 % (512-px light fields) of the HBM rate.
constexpr int kShiftThreads = 256;

__global__ void __launch_bounds__(kShiftThreads) lf_shift_kernel(const ShiftParams p, int rows_per_cta, int bands) {
  extern __shared__ __align__(16) float rows[];   // [rows_per_cta + 1][W]
  const int band = blockIdx.x % bands;
  int pl = blockIdx.x / bands;                  // plane index over (stack, batch, view, colour)
  const int planes_per_stack = p.batch * p.n * 3;
  const int stack = pl / planes_per_stack;
  pl -= stack * planes_per_stack;
  const int view = (pl / 3) % p.n;
  const int64_t plane_off = static_cast<int64_t>(pl) * p.H * p.W;
  const float* __restrict__ src = p.src[stack] + plane_off;
  float* __restrict__ dst = p.dst[stack] + plane_off;
  const float w0 = p.taps.w0[view], w1 = p.taps.w1[view];
  const int s0 = p.taps.s0[view], s1 = p.taps.s1[view];
  const bool has_w = stack != 1;                // h, i, d are resampled along W
  const bool has_v = stack != 0;                // v, i, d along H; the i stack with the opposite sign
  const int vsign = stack == 2 ? -1 : +1;
  const int W = p.W, H = p.H;
  const int y0 = band * rows_per_cta;
  const int ny = min(rows_per_cta, H - y0);
  // staged source rows: row (base + k) mod H at rows[k], k in [0, nstage).  The taps of output row y are the rows
  // src_index(y, s0) and src_index(y, s1); s1 = s0 +- 1, so over a band they cover one circular run.
  int base = y0, nstage = ny, k0_off = 0, k1_off = 0;
  if (has_v) {
    const int a0 = src_index(y0, s0, H, vsign), a1 = src_index(y0, s1, H, vsign);
    // circular order of the two taps: the one that comes first is the base
    const int d01 = ((a1 - a0) % H + H) % H;      // distance from tap 0 to tap 1 going up
    if (d01 <= 1) {
      base = a0; k0_off = 0; k1_off = d01;
    } else {                                        // tap 1 is one row below tap 0 (d01 == H - 1), or |s| >= H quirks
      base = a1; k1_off = 0; k0_off = ((a0 - a1) % H + H) % H;
    }
    nstage = ny + max(k0_off, k1_off);
    if (k0_off > 1 || k1_off > 1 || nstage > rows_per_cta + 1) {   // taps further apart than one row (tiny H): generic path
      nstage = -1;
    }
  }
  if (nstage > 0) {
    if ((W & 3) == 0) {
      const int w4 = W >> 2, total = nstage * w4;
#pragma unroll 4
      for (int i = threadIdx.x; i < total; i += kShiftThreads) {
        const int k = i / w4, x4 = i - k * w4;
        int r = base + k;
        if (r >= H) r -= H;
        reinterpret_cast<float4*>(rows)[i] = __ldg(reinterpret_cast<const float4*>(src + static_cast<int64_t>(r) * W) + x4);
      }
    } else {
      const int total = nstage * W;
      for (int i = threadIdx.x; i < total; i += kShiftThreads) {
        const int k = i / W, x = i - k * W;
        int r = base + k;
        if (r >= H) r -= H;
        rows[i] = __ldg(src + static_cast<int64_t>(r) * W + x);
      }
    }
  }
  __syncthreads();
  const int wq = (W + 3) >> 2;
  for (int i = threadIdx.x; i < ny * wq; i += kShiftThreads) {
    const int ky = i / wq, x0 = (i - ky * wq) * 4;
    const int y = y0 + ky;
    float o[4];
    if (nstage > 0) {
      const float* ra = rows + (has_v ? ky + k0_off : ky) * W;
      const float* rb = rows + (has_v ? ky + k1_off : ky) * W;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int x = x0 + j;
        if (x >= W) break;
        float t0, t1 = 0.f;
        if (has_w) {
          const int c0 = src_index(x, s0, W, +1), c1 = src_index(x, s1, W, +1);
          t0 = lerp2(ra[c0], w0, ra[c1], w1);
          if (has_v) t1 = lerp2(rb[c0], w0, rb[c1], w1);
        } else {
          t0 = ra[x];
          t1 = rb[x];
        }
        o[j] = has_v ? lerp2(t0, w0, t1, w1) : t0;
      }
    } else {                                        // generic path straight from global memory
      const float* ra = src + static_cast<int64_t>(src_index(y, s0, H, vsign)) * W;
      const float* rb = src + static_cast<int64_t>(src_index(y, s1, H, vsign)) * W;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int x = x0 + j;
        if (x >= W) break;
        const int c0 = has_w ? src_index(x, s0, W, +1) : x, c1 = has_w ? src_index(x, s1, W, +1) : x;
        const float t0 = has_w ? lerp2(__ldg(ra + c0), w0, __ldg(ra + c1), w1) : __ldg(ra + x);
        const float t1 = has_w ? lerp2(__ldg(rb + c0), w0, __ldg(rb + c1), w1) : __ldg(rb + x);
        o[j] = lerp2(t0, w0, t1, w1);
      }
    }
    float* d = dst + static_cast<int64_t>(y) * W + x0;
    if ((W & 3) == 0) {
      *reinterpret_cast<float4*>(d) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
      for (int j = 0; j < 4 && x0 + j < W; ++j) d[j] = o[j];
    }
  }
}

// Register-only variant for W % 4 == 0 (every real light field): one thread = 4 consecutive pixels x R consecutive rows
// of one plane.  The two W taps of the four pixels are 5 circularly consecutive source pixels
// (s1 = s0 +- 1), i.e. they lie inside two aligned float4 of the source row, so every load is an aligned, coalesced
// 128-bit load, the lane/row overlaps are served by L1/L2 and HBM sees each byte once.  No shared memory, no barriers.

__device__ __forceinline__ int eff_shift(int s, int n) { return (s == 0 || s >= n || -s >= n) ? 0 : s; }
__device__ __forceinline__ int wrap(int r, int n) {          // r in (-n, 2n)
  if (r < 0) r += n;
  if (r >= n) r -= n;
  return r;
}

// the 8 floats [a, a + 8) of a source row (a % 4 == 0, taken modulo W; W % 4 == 0 keeps each float4 on one side of the wrap)
__device__ __forceinline__ void load8(const float* __restrict__ row, int a, int W, float (&f)[8]) {
  const float4 lo = __ldg(reinterpret_cast<const float4*>(row + a));
  int b = a + 4;
  if (b >= W) b -= W;
  const float4 hi = __ldg(reinterpret_cast<const float4*>(row + b));
  f[0] = lo.x; f[1] = lo.y; f[2] = lo.z; f[3] = lo.w;
  f[4] = hi.x; f[5] = hi.y; f[6] = hi.z; f[7] = hi.w;
}

// W lerp of 4 pixels from the 8-float window: taps at window offsets o + d0 + j and o + d1 + j
__device__ __forceinline__ void wlerp4(const float (&f)[8], int o, int d0, int d1, float w0, float w1, float (&t)[4]) {
  float g[5];
  switch (o) {                                    // o is uniform per plane
    case 0: g[0] = f[0]; g[1] = f[1]; g[2] = f[2]; g[3] = f[3]; g[4] = f[4]; break;
    case 1: g[0] = f[1]; g[1] = f[2]; g[2] = f[3]; g[3] = f[4]; g[4] = f[5]; break;
    case 2: g[0] = f[2]; g[1] = f[3]; g[2] = f[4]; g[3] = f[5]; g[4] = f[6]; break;
    default: g[0] = f[3]; g[1] = f[4]; g[2] = f[5]; g[3] = f[6]; g[4] = f[7]; break;
  }
  if (d0 == 0) {
    // (a select, not g[j + d1]: a run-time index would move the window to local memory)
#pragma unroll
    for (int j = 0; j < 4; ++j) t[j] = lerp2(g[j], w0, d1 ? g[j + 1] : g[j], w1);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) t[j] = lerp2(g[j + 1], w0, g[j], w1);
  }
}

// W lerp of the packing kernel, taps ordered by column: window[O + j] * wa + window[O + j + DD] * wb (the float add commutes
// bit for bit, so this equals tap0 * w0 + tap1 * w1 of wlerp4).  Window offset and tap distance are template parameters:
// the selection costs no instructions
template <int V> struct IntTag { static constexpr int value = V; };
template <int O, int DD>
__device__ __forceinline__ void wlerp4s(const float (&f)[8], float wa, float wb, float (&t)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) t[j] = __fadd_rn(__fmul_rn(f[O + j], wa), __fmul_rn(f[O + j + DD], wb));
}

template <int R>
__global__ void __launch_bounds__(256) lf_shift_vec_kernel(const ShiftParams p, int planes) {
  // grid: x = blocks of 256 threads inside a plane, (y, z) = plane; the plane decode is block-uniform and the in-plane
  // index needs one 32-bit division (the flat 64-bit index cost ~300 of the kernel's 680 instructions per thread)
  const int W = p.W, H = p.H, w4 = W >> 2;
  const int nyb = (H + R - 1) / R;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int pl = blockIdx.z * gridDim.y + blockIdx.y;
  if (i >= nyb * w4 || pl >= planes) return;
  const int yb = i / w4;
  const int x0 = (i - yb * w4) * 4;
  const int planes_per_stack = p.batch * p.n * 3;
  const int stack = pl / planes_per_stack;
  pl -= stack * planes_per_stack;
  const int view = (pl / 3) % p.n;
  const int64_t plane_off = static_cast<int64_t>(pl) * H * W;
  const float* __restrict__ src = p.src[stack] + plane_off;
  float* __restrict__ dst = p.dst[stack] + plane_off;
  const float w0 = p.taps.w0[view], w1 = p.taps.w1[view];
  const int s0 = p.taps.s0[view], s1 = p.taps.s1[view];
  const bool has_w = stack != 1, has_v = stack != 0;
  const int vsign = stack == 2 ? -1 : +1;
  // W taps: source columns (x - e0) mod W and (x - e1) mod W; window start = the circularly smaller of the two
  const int e0 = eff_shift(s0, W), e1 = eff_shift(s1, W);
  int d0 = 0, d1 = 0, m = x0, o = 0, a = x0;
  if (has_w) {
    const int c0 = wrap(x0 - e0, W), c1 = wrap(x0 - e1, W);
    const int up = wrap(c1 - c0, W);              // 0: same column, 1: tap 1 right of tap 0, W - 1: left of it
    if (up <= 1) { m = c0; d0 = 0; d1 = up; } else { m = c1; d0 = 1; d1 = 0; }   // (other distances never reach here)
    o = m & 3;
    a = m - o;
  }
  const int y0 = yb * R;
  // one output row: all of its loads are issued before the arithmetic; with the row loop fully unrolled (no early exit)
  // the compiler interleaves the loads of all R rows
  auto do_row = [&](int y) {
    const int ra = has_v ? src_index(y, s0, H, vsign) : y;
    const int rb = has_v ? src_index(y, s1, H, vsign) : y;
    float t0[4], t1[4];
    if (has_w) {
      float f[8], g[8];
      load8(src + static_cast<int64_t>(ra) * W, a, W, f);
      if (has_v) load8(src + static_cast<int64_t>(rb) * W, a, W, g);
      wlerp4(f, o, d0, d1, w0, w1, t0);
      if (has_v) wlerp4(g, o, d0, d1, w0, w1, t1);
    } else {
      const float4 u = __ldg(reinterpret_cast<const float4*>(src + static_cast<int64_t>(ra) * W + x0));
      const float4 v = __ldg(reinterpret_cast<const float4*>(src + static_cast<int64_t>(rb) * W + x0));
      t0[0] = u.x; t0[1] = u.y; t0[2] = u.z; t0[3] = u.w;
      t1[0] = v.x; t1[1] = v.y; t1[2] = v.z; t1[3] = v.w;
    }
    float4 out;
    if (has_v) {
      out = make_float4(lerp2(t0[0], w0, t1[0], w1), lerp2(t0[1], w0, t1[1], w1), lerp2(t0[2], w0, t1[2], w1),
                        lerp2(t0[3], w0, t1[3], w1));
    } else {
      out = make_float4(t0[0], t0[1], t0[2], t0[3]);
    }
    __stcs(reinterpret_cast<float4*>(dst + static_cast<int64_t>(y) * W + x0), out);
  };
  if (y0 + R <= H) {
#pragma unroll
    for (int ky = 0; ky < R; ++ky) do_row(y0 + ky);
  } else {
    for (int y = y0; y < H; ++y) do_row(y);
  }
}

// ------------------------------------------------------------------------------------------------ pack
// views (B, C, H, W) f32 -> bf16 slots [B*(H+1)*(W+1)][ld]: channel-last, pixel (y, x) at slot (y+1, x+1), zero halo.
// One CTA = one slot row (b, sy) x 32 slot columns; transposes through shared memory so that both the f32 reads
// (along W) and the bf16 writes (along channels, 2*ld bytes per slot, consecutive slots contiguous) are coalesced.
// With `shift != nullptr` the value written is the Shift-resampled one (fused ESE path).
struct PackParams {
  const float* views;
  __nv_bfloat16* out;
  // multi-stack launches (blockIdx.z = stack): all stacks of a forward pass in ONE launch, optionally with a second copy
  // in another 16-bit format written from the same tile (the bf16 twin the weight-gradient GEMM reads)
  const float* views4[4];
  __nv_bfloat16* out4[4];
  __nv_bfloat16* out2_4[4];
  int stack4[4];
  int n_stacks, dtype2;
  int B, C, H, W, ld;
  int cw;                        // channels written per slot (multiple of 8, <= ld); columns [cw, ld) are left alone
  int residual;                  // write fp16(x - float(fp16(x))) instead of fp16(x) (split-precision lo block)
  int rows_per_cta;              // vector kernel: consecutive slot rows per CTA
  int do_shift, stack, n, dtype;
  ShiftTaps taps;
};

__device__ __forceinline__ float shifted_value(const float* __restrict__ plane, int y, int x, int H, int W, int stack,
                                               float w0, float w1, int s0, int s1) {
  const bool has_w = stack != 1, has_v = stack != 0;
  const int vsign = stack == 2 ? -1 : +1;
  const int r0 = has_v ? src_index(y, s0, H, vsign) : y;
  const int r1 = has_v ? src_index(y, s1, H, vsign) : y;
  const int c0 = has_w ? src_index(x, s0, W, +1) : x;
  const int c1 = has_w ? src_index(x, s1, W, +1) : x;
  const float* a = plane + static_cast<int64_t>(r0) * W;
  const float* b = plane + static_cast<int64_t>(r1) * W;
  const float t0 = has_w ? lerp2(__ldg(a + c0), w0, __ldg(a + c1), w1) : __ldg(a + x);
  if (!has_v) return t0;
  const float t1 = has_w ? lerp2(__ldg(b + c0), w0, __ldg(b + c1), w1) : __ldg(b + x);
  return lerp2(t0, w0, t1, w1);
}

constexpr int kPackTile = 128;                     // slot columns per CTA

__device__ __forceinline__ uint32_t pack_pair(float a, float b, const PackParams& p, int dtype) {
  if (p.residual) {
    a -= from16(to16(a, kFP16), kFP16);
    b -= from16(to16(b, kFP16), kFP16);
  }
  return pack16x2(a, b, dtype);
}
__device__ __forceinline__ uint32_t pack_pair(float a, float b, const PackParams& p) { return pack_pair(a, b, p, p.dtype); }

#define MMLF_PACK_SELECT()                                                                                           \
  const int zz = blockIdx.z;                                                                                          \
  const bool multi = p.n_stacks > 0;                                                                                  \
  const float* pviews = !multi ? p.views : zz == 0 ? p.views4[0] : zz == 1 ? p.views4[1] : zz == 2 ? p.views4[2] : p.views4[3]; \
  __nv_bfloat16* pout = !multi ? p.out : zz == 0 ? p.out4[0] : zz == 1 ? p.out4[1] : zz == 2 ? p.out4[2] : p.out4[3];    \
  __nv_bfloat16* out2 = !multi ? nullptr : zz == 0 ? p.out2_4[0] : zz == 1 ? p.out2_4[1] : zz == 2 ? p.out2_4[2] : p.out2_4[3]; \
  const int pstack = !multi ? p.stack : zz == 0 ? p.stack4[0] : zz == 1 ? p.stack4[1] : zz == 2 ? p.stack4[2] : p.stack4[3];

__global__ void __launch_bounds__(256) pack_views_kernel(const PackParams p) {
  __shared__ float tile[32][kPackTile + 1];        // [channel][slot column], 32 channels per pass
  MMLF_PACK_SELECT()
  const int Wp = p.W + 1, Hp = p.H + 1;
  const int sx0 = blockIdx.x * kPackTile;
  const int sy = blockIdx.y % Hp, b = blockIdx.y / Hp;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int64_t slot_row = (static_cast<int64_t>(b) * Hp + sy) * Wp;
  for (int cbase = 0; cbase < p.cw; cbase += 32) {
    // load: warp w handles channels w, w+8, ...; lanes run along x, four columns (32 apart) per lane in flight
    for (int c = wrp; c < 32; c += 8) {
      const int ch = cbase + c;
      const bool row_ok = ch < p.C && sy >= 1;
      const float* plane = pviews + (static_cast<int64_t>(b) * p.C + (row_ok ? ch : 0)) * p.H * p.W;
      float w0 = 0.f, w1 = 0.f;
      int s0 = 0, s1 = 0;
      if (p.do_shift && row_ok) {
        const int view = ch / 3;
        w0 = p.taps.w0[view]; w1 = p.taps.w1[view]; s0 = p.taps.s0[view]; s1 = p.taps.s1[view];
      }
      float v[kPackTile / 32];
#pragma unroll
      for (int j = 0; j < kPackTile / 32; ++j) {
        const int sx = sx0 + lane + 32 * j;
        v[j] = 0.f;
        if (row_ok && sx >= 1 && sx < Wp) {
          if (p.do_shift) v[j] = shifted_value(plane, sy - 1, sx - 1, p.H, p.W, pstack, w0, w1, s0, s1);
          else v[j] = __ldg(plane + static_cast<int64_t>(sy - 1) * p.W + (sx - 1));
        }
      }
#pragma unroll
      for (int j = 0; j < kPackTile / 32; ++j) tile[c][lane + 32 * j] = v[j];
    }
    __syncthreads();
    // store: each thread writes 8 consecutive channels (16 bytes) of one slot; 4 threads cover 32 channels
    const int cq = (threadIdx.x & 3) * 8;
#pragma unroll
    for (int pass = 0; pass < kPackTile / 64; ++pass) {
      const int slot = (threadIdx.x >> 2) + 64 * pass;
      const int sx = sx0 + slot;
      if (sx < Wp && cbase + cq < p.cw) {
        uint4 o;
        o.x = pack_pair(tile[cq][slot], tile[cq + 1][slot], p);
        o.y = pack_pair(tile[cq + 2][slot], tile[cq + 3][slot], p);
        o.z = pack_pair(tile[cq + 4][slot], tile[cq + 5][slot], p);
        o.w = pack_pair(tile[cq + 6][slot], tile[cq + 7][slot], p);
        *reinterpret_cast<uint4*>(pout + (slot_row + sx) * p.ld + cbase + cq) = o;
        if (out2) {
          o.x = pack_pair(tile[cq][slot], tile[cq + 1][slot], p, p.dtype2);
          o.y = pack_pair(tile[cq + 2][slot], tile[cq + 3][slot], p, p.dtype2);
          o.z = pack_pair(tile[cq + 4][slot], tile[cq + 5][slot], p, p.dtype2);
          o.w = pack_pair(tile[cq + 6][slot], tile[cq + 7][slot], p, p.dtype2);
          *reinterpret_cast<uint4*>(out2 + (slot_row + sx) * p.ld + cbase + cq) = o;
        }
      }
    }
    __syncthreads();
  }
}

// Vector variant for W % 4 == 0: one CTA = `rows_per_cta` consecutive slot rows x 128 pixels; a lane owns 4 consecutive
// pixels of a channel, read with aligned 128-bit loads (the Shift taps through the same two-float4 window as
// lf_shift_vec_kernel), transposed through shared memory and written as 16-byte channel groups.  Slot column 0 (the halo)
// is written by the first tile.  MODE (block-uniform): 0 = plain copy, 1 = W taps only (stack 0), 2 = H taps only
// (stack 1), 3 = both (stacks 2, 3).  The loads of a batch of channels (4; 2 with both taps = 8 x 128 bit) are all issued
// before the first dependent instruction: ncu had the one-channel-at-a-time loop waiting on a single load per thread
// (42 % of the stall samples on its first use).  Shared-memory columns are XOR-swizzled by the 8-channel group so that
// the channel-major reads of the store phase hit 32 different banks (they were 4-way conflicts: 132-float rows put the
// four channel groups of a slot on the same bank).
__device__ __forceinline__ float4 ldg4_if(const float* ptr, bool ok) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ok) v = __ldg(reinterpret_cast<const float4*>(ptr));
  return v;
}
__device__ __forceinline__ int pack_col(int c, int px) { return px ^ (((c >> 3) & 3) << 3); }

template <int MODE>
__device__ __forceinline__ void pack_vec_rows(const PackParams& p, const float* __restrict__ pviews, __nv_bfloat16* pout,
                                              __nv_bfloat16* out2, int pstack, float (*tile)[kPackTile + 4], int row0,
                                              int row1) {
  constexpr bool has_w = MODE == 1 || MODE == 3, has_v = MODE == 2 || MODE == 3;
  constexpr int kWarps = MODE == 0 ? 8 : 9;        // = blockDim.x / 32 (see pack_views_vec_kernel)
  const int Wp = p.W + 1, Hp = p.H + 1, W = p.W, H = p.H;
  const int X0 = blockIdx.x * kPackTile;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int x0 = X0 + 4 * lane;
  const bool x_ok = x0 < W;
  const int xl = x_ok ? x0 : 0;                    // addresses stay in range for the predicated-off lanes
  const int vsign = pstack == 2 ? -1 : +1;
  const int cq = (threadIdx.x & 3) * 8;
  for (int row = row0; row < row1; ++row) {
    const int sy = row % Hp, b = row / Hp;
    const int y = sy >= 1 ? sy - 1 : 0;
    const int64_t slot_row = static_cast<int64_t>(row) * Wp;
    for (int cbase = 0; cbase < p.cw; cbase += 32) {
      if (MODE == 0) {
        // plain copy: warp w owns channels w, w + 8, w + 16, w + 24 of the group; the four loads are issued together
        float4 val[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int ch = cbase + wrp + 8 * j;
          const int chl = ch < p.C ? ch : 0;
          const float* plane = pviews + (static_cast<int64_t>(b) * p.C + chl) * H * W;
          val[j] = ldg4_if(plane + static_cast<int64_t>(y) * W + xl, ch < p.C && sy >= 1 && x_ok);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = wrp + 8 * j;
          *reinterpret_cast<float4*>(&tile[c][pack_col(c, 4 * lane)]) = val[j];
        }
      } else {
        // Shift fused in: warp w owns VIEWS (three colour planes with the same taps), so the tap rows, the window start
        // and the weights are derived once per view instead of once per channel (the kernel is instruction-issue bound:
        // ~200 instructions per channel and lane group before, most of them index arithmetic)
        const int cend = cbase + 32 < p.C ? cbase + 32 : p.C;
        for (int view = cbase / 3 + wrp; 3 * view < cend; view += kWarps) {
          const float w0 = p.taps.w0[view], w1 = p.taps.w1[view];
          const int s0 = p.taps.s0[view], s1 = p.taps.s1[view];
          int ra = y, rb = y;
          if (has_v) { ra = src_index(y, s0, H, vsign); rb = src_index(y, s1, H, vsign); }
          int o = 0, dd = 0, a = xl, a2 = xl;
          float wa = w0, wb = w1;
          if (has_w) {
            // source columns (x - e0) mod W and (x - e1) mod W: neighbours (or equal); the 8-float window starts at the
            // aligned float4 holding the circularly smaller one.  tap0 * w0 + tap1 * w1 is evaluated as
            // window[j] * wa + window[j + dd] * wb (the float add commutes bit for bit)
            const int e0 = eff_shift(s0, W), e1 = eff_shift(s1, W);
            const int c0 = wrap(xl - e0, W), c1 = wrap(xl - e1, W);
            const int up = wrap(c1 - c0, W);
            int m;
            if (up <= 1) { m = c0; dd = up; } else { m = c1; dd = 1; wa = w1; wb = w0; }
            o = m & 3;
            a = m - o;
            a2 = a + 4;
            if (a2 >= W) a2 -= W;
          }
          // every address below is valid memory (predicated-off lanes / halo rows use x = 0 / y = 0, C is a multiple of
          // 3), so the loads are unconditional and the zero halo is a select at the end: no zero-initialisation and no
          // predicate logic around the 12 loads of a view
          const float* vbase = pviews + (static_cast<int64_t>(b) * p.C + 3 * view) * H * W;
          const int HW = H * W;
          const int oa = ra * W + a, oa2 = ra * W + a2, ob = rb * W + a, ob2 = rb * W + a2;
          const bool zero = !(sy >= 1 && x_ok);
          float fa[3][8], fb[3][8];
          bool live[3];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int ch = 3 * view + k;
            live[k] = ch >= cbase && ch < cend;
            const float* plane = vbase + k * HW;
            const float4 lo = __ldg(reinterpret_cast<const float4*>(plane + oa));
            fa[k][0] = lo.x; fa[k][1] = lo.y; fa[k][2] = lo.z; fa[k][3] = lo.w;
            if (has_w) {
              const float4 hi = __ldg(reinterpret_cast<const float4*>(plane + oa2));
              fa[k][4] = hi.x; fa[k][5] = hi.y; fa[k][6] = hi.z; fa[k][7] = hi.w;
            }
            if (has_v) {
              const float4 lo2 = __ldg(reinterpret_cast<const float4*>(plane + ob));
              fb[k][0] = lo2.x; fb[k][1] = lo2.y; fb[k][2] = lo2.z; fb[k][3] = lo2.w;
              if (has_w) {
                const float4 hi2 = __ldg(reinterpret_cast<const float4*>(plane + ob2));
                fb[k][4] = hi2.x; fb[k][5] = hi2.y; fb[k][6] = hi2.z; fb[k][7] = hi2.w;
              }
            }
          }
          // arithmetic of the three colour planes, specialised on the (view-uniform) window offset and tap distance
          auto finish = [&](auto tag) {
            constexpr int O = decltype(tag)::value >> 1, DD = decltype(tag)::value & 1;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              if (!live[k]) continue;
              const int c = 3 * view + k - cbase;
              float t0[4], t1[4];
              if (has_w) {
                wlerp4s<O, DD>(fa[k], wa, wb, t0);
                if (has_v) wlerp4s<O, DD>(fb[k], wa, wb, t1);
              } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) { t0[q] = fa[k][q]; if (has_v) t1[q] = fb[k][q]; }
              }
              float4 val;
              if (has_v) val = make_float4(lerp2(t0[0], w0, t1[0], w1), lerp2(t0[1], w0, t1[1], w1),
                                           lerp2(t0[2], w0, t1[2], w1), lerp2(t0[3], w0, t1[3], w1));
              else val = make_float4(t0[0], t0[1], t0[2], t0[3]);
              if (zero) val = make_float4(0.f, 0.f, 0.f, 0.f);
              *reinterpret_cast<float4*>(&tile[c][pack_col(c, 4 * lane)]) = val;
            }
          };
          if (!has_w) {
            finish(IntTag<0>{});
          } else {
            switch (2 * o + dd) {                  // uniform across the warp
              case 0: finish(IntTag<0>{}); break;
              case 1: finish(IntTag<1>{}); break;
              case 2: finish(IntTag<2>{}); break;
              case 3: finish(IntTag<3>{}); break;
              case 4: finish(IntTag<4>{}); break;
              case 5: finish(IntTag<5>{}); break;
              case 6: finish(IntTag<6>{}); break;
              default: finish(IntTag<7>{}); break;
            }
          }
        }
        // channels of the group beyond C: zeros
        for (int c = (p.C > cbase ? p.C - cbase : 0) + wrp; c < 32; c += kWarps)
          *reinterpret_cast<float4*>(&tile[c][pack_col(c, 4 * lane)]) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncthreads();
#pragma unroll
      for (int pass = 0; pass < kPackTile / 64; ++pass) {
        const int px = (threadIdx.x >> 2) + 64 * pass;
        if (threadIdx.x < 256 && X0 + px < W && cbase + cq < p.cw) {
          const int col = pack_col(cq, px);        // the 8 channels of a group share the swizzle
          float v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = tile[cq + k][col];
          uint4 q;
          q.x = pack_pair(v[0], v[1], p); q.y = pack_pair(v[2], v[3], p);
          q.z = pack_pair(v[4], v[5], p); q.w = pack_pair(v[6], v[7], p);
          *reinterpret_cast<uint4*>(pout + (slot_row + X0 + px + 1) * p.ld + cbase + cq) = q;
          if (out2) {
            q.x = pack_pair(v[0], v[1], p, p.dtype2); q.y = pack_pair(v[2], v[3], p, p.dtype2);
            q.z = pack_pair(v[4], v[5], p, p.dtype2); q.w = pack_pair(v[6], v[7], p, p.dtype2);
            *reinterpret_cast<uint4*>(out2 + (slot_row + X0 + px + 1) * p.ld + cbase + cq) = q;
          }
        }
      }
      __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x * 8 < p.cw) {          // halo column sx = 0
      *reinterpret_cast<uint4*>(pout + slot_row * p.ld + threadIdx.x * 8) = make_uint4(0u, 0u, 0u, 0u);
      if (out2) *reinterpret_cast<uint4*>(out2 + slot_row * p.ld + threadIdx.x * 8) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

// SHIFT = false: 256 threads, warp w loads channels w, w + 8, ...; SHIFT = true: 288 threads, one warp per view of a
// 9-view stack (ncu: with 8 warps the ninth view made warp 0 work twice as long and the other seven wait at the barrier,
// 27 % of the stall samples); the ninth warp sits out the store phase.
template <bool SHIFT>
__global__ void __launch_bounds__(SHIFT ? 288 : 256, SHIFT ? 3 : 4) pack_views_vec_kernel(const PackParams p) {
  __shared__ __align__(16) float tile[32][kPackTile + 4];   // [channel][pixel ^ swizzle], 16-byte aligned rows
  MMLF_PACK_SELECT()
  const int n_rows = p.B * (p.H + 1);
  const int row0 = blockIdx.y * p.rows_per_cta;
  const int row1 = row0 + p.rows_per_cta < n_rows ? row0 + p.rows_per_cta : n_rows;
  if (!SHIFT) pack_vec_rows<0>(p, pviews, pout, out2, pstack, tile, row0, row1);
  else if (pstack == 0) pack_vec_rows<1>(p, pviews, pout, out2, pstack, tile, row0, row1);
  else if (pstack == 1) pack_vec_rows<2>(p, pviews, pout, out2, pstack, tile, row0, row1);
  else pack_vec_rows<3>(p, pviews, pout, out2, pstack, tile, row0, row1);
}

// ------------------------------------------------------------------------------------------------ texture mask
// create_mask_texture (hci4d.py:38-69): mean over 3 colours x wsize^2 zero-padded neighbours of |neighbour - centre|,
// thresholded, with a margin of wsize / 2 masked out.  The reference materialises the unfolded (1, 3 * wsize^2, H * W)
// tensor (1587 x the image for wsize 23); here one CTA stages a (32 + 2r) x (32 + 2r) x 3 tile in shared memory and
// every thread walks its pixel's window from there.
constexpr int kTexTile = 32;

__global__ void __launch_bounds__(256) texture_mask_kernel(const float* __restrict__ center, int B, int H, int W,
                                                            int wsize, float threshold, int32_t* __restrict__ mask,
                                                            float* __restrict__ mae) {
  extern __shared__ float tex_tile[];               // [3][T][T], T = kTexTile + 2r
  const int r = wsize / 2, T = kTexTile + 2 * r;
  const int b = blockIdx.z, y0 = blockIdx.y * kTexTile, x0 = blockIdx.x * kTexTile;
  const float* img = center + static_cast<int64_t>(b) * 3 * H * W;
  for (int i = threadIdx.x; i < 3 * T * T; i += blockDim.x) {
    const int c = i / (T * T), rem = i - c * T * T;
    const int ty = rem / T, tx = rem - ty * T;
    const int y = y0 + ty - r, x = x0 + tx - r;
    tex_tile[i] = (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(img + (static_cast<int64_t>(c) * H + y) * W + x) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kTexTile * kTexTile; i += blockDim.x) {
    const int py = i / kTexTile, px = i - py * kTexTile;
    const int y = y0 + py, x = x0 + px;
    if (y >= H || x >= W) continue;
    float acc = 0.f;
    for (int c = 0; c < 3; ++c) {
      const float* t = tex_tile + c * T * T;
      const float ctr = t[(py + r) * T + px + r];
      for (int dy = 0; dy < wsize; ++dy) {
        const float* row = t + (py + dy) * T + px;
        float racc = 0.f;
        for (int dx = 0; dx < wsize; ++dx) racc += fabsf(row[dx] - ctr);
        acc += racc;
      }
    }
    const float m = acc / static_cast<float>(3 * wsize * wsize);
    const bool inside = y >= r && y < H - r && x >= r && x < W - r;
    const int64_t o = (static_cast<int64_t>(b) * H + y) * W + x;
    mask[o] = (m >= threshold && inside) ? 1 : 0;
    if (mae) mae[o] = m;
  }
}

}  // namespace mmlf

using namespace mmlf;

extern "C" int mmlf_shift_taps(double disp, int n, float* w0, float* w1, int* s0, int* s1) {
  MMLF_REQUIRE(n >= 1 && n <= 16, "shift_taps: n must be in [1, 16]");
  ShiftTaps t;
  host_taps(disp, n, t);
  for (int i = 0; i < n; ++i) {
    w0[i] = t.w0[i];
    w1[i] = t.w1[i];
    s0[i] = t.s0[i];
    s1[i] = t.s1[i];
  }
  return 0;
}

extern "C" int mmlf_lf_extract_u8(const uint8_t* views, int n, int H, int W, float* h, float* v, float* i, float* d,
                                  float* center, void* stream) {
  MMLF_REQUIRE(views && h && v && i && d, "lf_extract: null buffer");
  MMLF_REQUIRE(n >= 1 && n <= 16 && (n & 1), "lf_extract: n must be odd and <= 16");
  MMLF_REQUIRE(W % 4 == 0, "lf_extract: W must be a multiple of 4");
  ExtractParams p;
  p.n = n; p.H = H; p.W = W; p.center = center;
  p.dst[0] = h; p.dst[1] = v; p.dst[2] = i; p.dst[3] = d;
  for (int k = 0; k < n; ++k) {
    p.idx[0][k] = (n / 2) * n + k;               // us  (hci4d.py:143)
    p.idx[1][k] = n / 2 + n * k;                 // vs  (hci4d.py:144)
    const int kk = n - 1 - k;                    // ids reversed (hci4d.py:147-148)
    p.idx[2][k] = n - kk - 1 + n * kk;
    p.idx[3][k] = k + n * k;                     // dds (hci4d.py:149)
  }
  const int64_t per_view = static_cast<int64_t>(H) * (W / 4);
  dim3 grid(static_cast<unsigned>(ceil_div64(per_view, 256)), 4 * n);
  lf_extract_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(views, p);
  return check_launch("lf_extract_kernel");
}

extern "C" int mmlf_lf_shift(const float* src_h, const float* src_v, const float* src_i, const float* src_d,
                             float* dst_h, float* dst_v, float* dst_i, float* dst_d, int batch, int n, int H, int W,
                             double disp, void* stream) {
  MMLF_REQUIRE(src_h && src_v && src_i && src_d && dst_h && dst_v && dst_i && dst_d, "lf_shift: null buffer");
  MMLF_REQUIRE(n >= 1 && n <= 16, "lf_shift: n must be in [1, 16]");
  MMLF_REQUIRE(src_h != dst_h && src_v != dst_v && src_i != dst_i && src_d != dst_d, "lf_shift is out of place");
  ShiftParams p;
  p.src[0] = src_h; p.src[1] = src_v; p.src[2] = src_i; p.src[3] = src_d;
  p.dst[0] = dst_h; p.dst[1] = dst_v; p.dst[2] = dst_i; p.dst[3] = dst_d;
  p.batch = batch; p.n = n; p.H = H; p.W = W;
  host_taps(disp, n, p.taps);
  const int64_t planes = static_cast<int64_t>(4) * batch * n * 3;
  {
    // W taps must be circular neighbours for the register kernel (always true for W >= 3: s1 = s0 +- 1)
    bool vec_ok = (W % 4 == 0) && W >= 8;
    for (int s = 0; s < 4 && vec_ok; ++s)
      vec_ok = (reinterpret_cast<uintptr_t>(p.src[s]) % 16 == 0) && (reinterpret_cast<uintptr_t>(p.dst[s]) % 16 == 0);
    if (vec_ok) {
      static int rows_per_thread = -1;
      if (rows_per_thread < 0) {
        const char* e = getenv("MMLF_SHIFT_ROWS");
        rows_per_thread = e ? atoi(e) : 4;
      }
      const int R = rows_per_thread;
      const int64_t in_plane = static_cast<int64_t>(ceil_div(H, R)) * (W / 4);
      MMLF_REQUIRE(in_plane < (1ll << 30) && planes < (1ll << 30), "lf_shift: too many planes / pixels");
      const unsigned gy = planes < 32768 ? static_cast<unsigned>(planes) : 32768u;
      // block size with the fewest idle threads in a plane's last block (96-px patches: 576 threads = 3 x 192)
      unsigned bs = 256;
      int64_t waste = ceil_div64(in_plane, 256) * 256 - in_plane;
      for (unsigned c : {192u, 128u}) {
        const int64_t w = ceil_div64(in_plane, c) * c - in_plane;
        if (w < waste) { waste = w; bs = c; }
      }
      const dim3 grid(static_cast<unsigned>(ceil_div64(in_plane, bs)), gy, static_cast<unsigned>(ceil_div64(planes, gy)));
      const int np = static_cast<int>(planes);
      cudaStream_t st = static_cast<cudaStream_t>(stream);
      if (R == 1) lf_shift_vec_kernel<1><<<grid, bs, 0, st>>>(p, np);
      else if (R == 2) lf_shift_vec_kernel<2><<<grid, bs, 0, st>>>(p, np);
      else if (R == 8) lf_shift_vec_kernel<8><<<grid, bs, 0, st>>>(p, np);
      else lf_shift_vec_kernel<4><<<grid, bs, 0, st>>>(p, np);
      return check_launch("lf_shift_vec_kernel");
    }
  }
  // band height: >= 8192 elements (32 KB in, 32 KB out) per CTA, bands of equal height
  int rows = (8192 + W - 1) / W;
  if (rows > H) rows = H;
  const int bands = ceil_div(H, rows);
  rows = ceil_div(H, bands);
  const size_t smem = static_cast<size_t>(rows + 1) * W * sizeof(float);
  MMLF_REQUIRE(smem <= 200 * 1024, "lf_shift: rows of %d pixels are too wide for the shared-memory band", W);
  MMLF_REQUIRE(planes * bands < (1ll << 31), "lf_shift: too many planes");
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(lf_shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    MMLF_REQUIRE(e == cudaSuccess, "lf_shift: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    smem_set = 200 * 1024;
  }
  lf_shift_kernel<<<static_cast<unsigned>(planes * bands), kShiftThreads, smem, static_cast<cudaStream_t>(stream)>>>(p, rows, bands);
  return check_launch("lf_shift_kernel");
}

// rows per CTA of the vector kernel: enough CTAs for ~8 waves of 8 resident blocks per SM, at most 8 rows each
// (MMLF_PACK_ROWS overrides, for sweeps)
static int pack_rows_per_cta(int64_t n_rows, int tiles_x, int n_stacks) {
  const char* e = getenv("MMLF_PACK_ROWS");           // read per call: the tests switch it inside one process
  const int forced = e ? atoi(e) : 0;
  if (forced > 0) return forced < n_rows ? forced : static_cast<int>(n_rows);
  const int64_t target = 148ll * 8 * 8;
  int64_t r = n_rows * tiles_x * n_stacks / target;
  if (r < 1) r = 1;
  if (r > 8) r = 8;
  return static_cast<int>(r);
}

static int launch_pack(const float* views, int B, int C, int H, int W, void* out, int ld, int dtype, int do_shift,
                       int stack, int n, double disp, void* stream, int cw = 0, int residual = 0) {
  MMLF_REQUIRE(dtype == 0 || dtype == 1, "pack_views: dtype must be 0 (bf16) or 1 (fp16)");
  MMLF_REQUIRE(views && out, "pack_views: null buffer");
  MMLF_REQUIRE(ld % 8 == 0 && ld >= C, "pack_views: ld %d must be a multiple of 8 and >= C %d", ld, C);
  MMLF_REQUIRE(static_cast<int64_t>(B) * (H + 1) <= 65535 * 1024ll, "pack_views: batch too large");
  PackParams p;
  p.n_stacks = 0; p.dtype2 = dtype;
  p.views = views; p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.B = B; p.C = C; p.H = H; p.W = W; p.ld = ld;
  p.do_shift = do_shift; p.stack = stack; p.n = n; p.dtype = dtype;
  p.cw = cw ? cw : ld;
  p.residual = residual;
  p.rows_per_cta = 1;
  MMLF_REQUIRE(p.cw % 8 == 0 && p.cw >= C && p.cw <= ld, "pack_views: bad written-channel count %d", p.cw);
  if (do_shift) host_taps(disp, n, p.taps);
  const int64_t rows = static_cast<int64_t>(B) * (H + 1);
  // grid.y <= 65535: split the batch into slabs
  const int slab_b = static_cast<int>(65535 / (H + 1)) > 0 ? static_cast<int>(65535 / (H + 1)) : 1;
  (void)rows;
  for (int b0 = 0; b0 < B; b0 += slab_b) {
    PackParams q = p;
    const int nb = B - b0 < slab_b ? B - b0 : slab_b;
    q.B = nb;
    q.views = views + static_cast<int64_t>(b0) * C * H * W;
    q.out = p.out + static_cast<int64_t>(b0) * (H + 1) * (W + 1) * ld;
    const bool vec = W % 4 == 0 && W >= 8 && reinterpret_cast<uintptr_t>(q.views) % 16 == 0 &&
                     reinterpret_cast<uintptr_t>(q.out) % 16 == 0;
    if (vec) {
      q.rows_per_cta = pack_rows_per_cta(static_cast<int64_t>(nb) * (H + 1), ceil_div(W, kPackTile), 1);
      dim3 grid(ceil_div(W, kPackTile), static_cast<unsigned>(ceil_div(nb * (H + 1), q.rows_per_cta)));
      if (q.do_shift) pack_views_vec_kernel<true><<<grid, 288, 0, static_cast<cudaStream_t>(stream)>>>(q);
      else pack_views_vec_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(q);
    } else {
      dim3 grid(ceil_div(W + 1, kPackTile), static_cast<unsigned>(nb * (H + 1)));
      pack_views_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(q);
    }
    if (int rc = check_launch("pack_views_kernel")) return rc;
  }
  return 0;
}

extern "C" int mmlf_pack_views(const float* views, int B, int C, int H, int W, void* out, int ld, int dtype,
                               void* stream) {
  return launch_pack(views, B, C, H, W, out, ld, dtype, 0, 0, 0, 0.0, stream);
}

extern "C" int mmlf_pack_views_split(const float* views, int B, int C, int H, int W, void* out, int ld, int c_pad,
                                     void* stream) {
  MMLF_REQUIRE(c_pad % 8 == 0 && c_pad >= C && ld >= 2 * c_pad, "pack_views_split: bad c_pad %d / ld %d", c_pad, ld);
  if (int rc = launch_pack(views, B, C, H, W, out, ld, kFP16, 0, 0, 0, 0.0, stream, c_pad, 0)) return rc;
  return launch_pack(views, B, C, H, W, reinterpret_cast<uint16_t*>(out) + c_pad, ld, kFP16, 0, 0, 0, 0.0, stream, c_pad, 1);
}

extern "C" int mmlf_shift_pack(const float* src, int stack, int B, int n, int H, int W, double disp, void* out,
                               int ld, int dtype, void* stream) {
  MMLF_REQUIRE(stack >= 0 && stack < 4, "shift_pack: stack must be 0..3");
  MMLF_REQUIRE(n >= 1 && n <= 16, "shift_pack: n must be in [1, 16]");
  return launch_pack(src, B, n * 3, H, W, out, ld, dtype, 1, stack, n, disp, stream);
}

extern "C" int mmlf_pack_stacks(const float* const* views, const int* stacks, int n_stacks, int B, int n, int H, int W,
                                void* const* out, void* const* out2, int ld, int dtype, int dtype2, int do_shift,
                                double disp, void* stream) {
  MMLF_REQUIRE(views && stacks && out && n_stacks >= 1 && n_stacks <= 4, "pack_stacks: 1..4 stacks");
  MMLF_REQUIRE((dtype == 0 || dtype == 1) && (dtype2 == 0 || dtype2 == 1), "pack_stacks: dtype must be 0 (bf16) or 1 (fp16)");
  MMLF_REQUIRE(n >= 1 && n <= 16, "pack_stacks: n must be in [1, 16]");
  const int C = n * 3;
  MMLF_REQUIRE(ld % 8 == 0 && ld >= C, "pack_stacks: ld %d must be a multiple of 8 and >= C %d", ld, C);
  MMLF_REQUIRE(static_cast<int64_t>(B) * (H + 1) <= 65535, "pack_stacks: batch * (H + 1) must fit one grid dimension");
  PackParams p;
  p.views = nullptr; p.out = nullptr;
  p.n_stacks = n_stacks; p.dtype = dtype; p.dtype2 = dtype2;
  p.B = B; p.C = C; p.H = H; p.W = W; p.ld = ld; p.cw = ld; p.residual = 0; p.rows_per_cta = 1;
  p.do_shift = do_shift; p.stack = 0; p.n = n;
  bool vec = W % 4 == 0 && W >= 8;
  for (int i = 0; i < 4; ++i) {
    const int j = i < n_stacks ? i : 0;
    MMLF_REQUIRE(views[j] && out[j], "pack_stacks: null buffer");
    MMLF_REQUIRE(stacks[j] >= 0 && stacks[j] < 4, "pack_stacks: stack must be 0..3");
    p.views4[i] = views[j];
    p.out4[i] = reinterpret_cast<__nv_bfloat16*>(out[j]);
    p.out2_4[i] = out2 ? reinterpret_cast<__nv_bfloat16*>(out2[j]) : nullptr;
    p.stack4[i] = stacks[j];
    vec = vec && reinterpret_cast<uintptr_t>(views[j]) % 16 == 0 && reinterpret_cast<uintptr_t>(out[j]) % 16 == 0 &&
          (!out2 || reinterpret_cast<uintptr_t>(out2[j]) % 16 == 0);
  }
  if (do_shift) host_taps(disp, n, p.taps);
  if (vec) {
    p.rows_per_cta = pack_rows_per_cta(static_cast<int64_t>(B) * (H + 1), ceil_div(W, kPackTile), n_stacks);
    dim3 grid(ceil_div(W, kPackTile), static_cast<unsigned>(ceil_div(B * (H + 1), p.rows_per_cta)), n_stacks);
    if (p.do_shift) pack_views_vec_kernel<true><<<grid, 288, 0, static_cast<cudaStream_t>(stream)>>>(p);
    else pack_views_vec_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  } else {
    dim3 grid(ceil_div(W + 1, kPackTile), static_cast<unsigned>(B * (H + 1)), n_stacks);
    pack_views_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  }
  return check_launch("pack_stacks");
}

extern "C" int mmlf_texture_mask(const float* center, int B, int H, int W, int wsize, double threshold, int32_t* mask,
                                 float* mae, void* stream) {
  MMLF_REQUIRE(center && mask, "texture_mask: null buffer");
  MMLF_REQUIRE(wsize >= 1 && (wsize & 1) && wsize <= 63, "texture_mask: wsize must be odd and <= 63");
  MMLF_REQUIRE(B >= 1 && B <= 65535 && H >= 1 && W >= 1, "texture_mask: bad shape");
  const int T = kTexTile + 2 * (wsize / 2);
  const size_t smem = static_cast<size_t>(3) * T * T * sizeof(float);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(texture_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    MMLF_REQUIRE(e == cudaSuccess, "texture_mask: %s", cudaGetErrorString(e));
    configured = smem;
  }
  dim3 grid(ceil_div(W, kTexTile), ceil_div(H, kTexTile), B);
  texture_mask_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(center, B, H, W, wsize,
                                                                              static_cast<float>(threshold), mask, mae);
  return check_launch("texture_mask_kernel");
}
