// Light-field input kernels: view-index extraction (u8 -> f32), the disparity Shift resampler and the
// conversions into the bf16 slot layout.  Bandwidth-bound; bit-exact with the reference
// (/root/reference/mmlf/data/hci4d.py:142-193, 907-990).
#include <math.h>

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
    v.x = __fdiv_rn(static_cast<float>(px[c]), 255.f);
    v.y = __fdiv_rn(static_cast<float>(px[3 + c]), 255.f);
    v.z = __fdiv_rn(static_cast<float>(px[6 + c]), 255.f);
    v.w = __fdiv_rn(static_cast<float>(px[9 + c]), 255.f);
    *reinterpret_cast<float4*>(dst + c * plane) = v;
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

// One CTA = one output row of one (stack, batch, view, colour) plane.  The one or two source rows are staged in
// shared memory with aligned 128-bit loads; the shifted (unaligned, wrapping) taps are then read from there.
__global__ void __launch_bounds__(128) lf_shift_kernel(const ShiftParams p) {
  extern __shared__ float rows[];               // [2][W]
  const int y = blockIdx.x;
  int pl = blockIdx.y;                          // plane index over (stack, batch, view, colour)
  const int planes_per_stack = p.batch * p.n * 3;
  const int stack = pl / planes_per_stack;
  pl -= stack * planes_per_stack;
  const int view = (pl / 3) % p.n;
  const int64_t plane_off = static_cast<int64_t>(pl) * p.H * p.W;
  const float* src = p.src[stack] + plane_off;
  float* dst = p.dst[stack] + plane_off;
  const float w0 = p.taps.w0[view], w1 = p.taps.w1[view];
  const int s0 = p.taps.s0[view], s1 = p.taps.s1[view];
  const bool has_w = stack != 1;                // h, i, d are resampled along W
  const bool has_v = stack != 0;                // v, i, d along H; the i stack with the opposite sign
  const int vsign = stack == 2 ? -1 : +1;
  const int r0 = has_v ? src_index(y, s0, p.H, vsign) : y;
  const int r1 = has_v ? src_index(y, s1, p.H, vsign) : y;
  const int W = p.W;
  if ((W & 3) == 0) {
    for (int x = threadIdx.x * 4; x < W; x += blockDim.x * 4) {
      *reinterpret_cast<float4*>(rows + x) = __ldg(reinterpret_cast<const float4*>(src + static_cast<int64_t>(r0) * W + x));
      if (has_v)
        *reinterpret_cast<float4*>(rows + W + x) = __ldg(reinterpret_cast<const float4*>(src + static_cast<int64_t>(r1) * W + x));
    }
  } else {
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
      rows[x] = __ldg(src + static_cast<int64_t>(r0) * W + x);
      if (has_v) rows[W + x] = __ldg(src + static_cast<int64_t>(r1) * W + x);
    }
  }
  __syncthreads();
  for (int x0 = threadIdx.x * 4; x0 < W; x0 += blockDim.x * 4) {
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int x = x0 + j;
      if (x >= W) break;
      float t0, t1 = 0.f;
      if (has_w) {
        const int c0 = src_index(x, s0, W, +1), c1 = src_index(x, s1, W, +1);
        t0 = lerp2(rows[c0], w0, rows[c1], w1);
        if (has_v) t1 = lerp2(rows[W + c0], w0, rows[W + c1], w1);
      } else {
        t0 = rows[x];
        t1 = rows[W + x];
      }
      o[j] = has_v ? lerp2(t0, w0, t1, w1) : t0;
    }
    float* d = dst + static_cast<int64_t>(y) * W + x0;
    if ((W & 3) == 0) {
      *reinterpret_cast<float4*>(d) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
      for (int j = 0; j < 4 && x0 + j < W; ++j) d[j] = o[j];
    }
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
  int B, C, H, W, ld;
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

__global__ void __launch_bounds__(256) pack_views_kernel(const PackParams p) {
  __shared__ float tile[32][33];                 // [channel][slot column], C <= 32 per pass
  const int Wp = p.W + 1, Hp = p.H + 1;
  const int sx0 = blockIdx.x * 32;
  const int sy = blockIdx.y % Hp, b = blockIdx.y / Hp;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int64_t slot_row = (static_cast<int64_t>(b) * Hp + sy) * Wp;
  for (int cbase = 0; cbase < p.ld; cbase += 32) {
    // load: warp w handles channels w, w+8, ...; lanes run along x
    for (int c = wrp; c < 32; c += 8) {
      const int ch = cbase + c;
      const int sx = sx0 + lane;
      float v = 0.f;
      if (ch < p.C && sy >= 1 && sx >= 1 && sx < Wp) {
        const float* plane = p.views + (static_cast<int64_t>(b) * p.C + ch) * p.H * p.W;
        if (p.do_shift) {
          const int view = ch / 3;
          v = shifted_value(plane, sy - 1, sx - 1, p.H, p.W, p.stack, p.taps.w0[view], p.taps.w1[view],
                            p.taps.s0[view], p.taps.s1[view]);
        } else {
          v = __ldg(plane + static_cast<int64_t>(sy - 1) * p.W + (sx - 1));
        }
      }
      tile[c][lane] = v;
    }
    __syncthreads();
    // store: each thread writes 4 consecutive channels (8 bytes) of one slot; 8 threads cover 32 channels
    const int slot = threadIdx.x >> 3, cq = (threadIdx.x & 7) * 4;
    const int sx = sx0 + slot;
    if (sx < Wp && cbase + cq < p.ld) {
      uint2 o;
      o.x = pack16x2(tile[cq][slot], tile[cq + 1][slot], p.dtype);
      o.y = pack16x2(tile[cq + 2][slot], tile[cq + 3][slot], p.dtype);
      *reinterpret_cast<uint2*>(p.out + (slot_row + sx) * p.ld + cbase + cq) = o;
    }
    __syncthreads();
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
  // grid.y is limited to 65535: launch in slabs of planes (same kernel, offset pointers)
  const int64_t per_stack = static_cast<int64_t>(batch) * n * 3;
  (void)planes;
  if (4 * per_stack <= 65535) {
    dim3 grid(H, static_cast<unsigned>(4 * per_stack));
    lf_shift_kernel<<<grid, 128, 2 * W * sizeof(float), static_cast<cudaStream_t>(stream)>>>(p);
    return check_launch("lf_shift_kernel");
  }
  // large batches: one launch per batch slab so that plane indices stay below the grid.y limit
  const int slab = static_cast<int>(65535 / (4 * n * 3));
  for (int b0 = 0; b0 < batch; b0 += slab) {
    ShiftParams q = p;
    const int nb = batch - b0 < slab ? batch - b0 : slab;
    const int64_t off = static_cast<int64_t>(b0) * n * 3 * H * W;
    for (int s = 0; s < 4; ++s) {
      q.src[s] = p.src[s] + off;
      q.dst[s] = p.dst[s] + off;
    }
    q.batch = nb;
    dim3 grid(H, static_cast<unsigned>(4 * nb * n * 3));
    lf_shift_kernel<<<grid, 128, 2 * W * sizeof(float), static_cast<cudaStream_t>(stream)>>>(q);
    if (int rc = check_launch("lf_shift_kernel")) return rc;
  }
  return 0;
}

static int launch_pack(const float* views, int B, int C, int H, int W, void* out, int ld, int dtype, int do_shift,
                       int stack, int n, double disp, void* stream) {
  MMLF_REQUIRE(dtype == 0 || dtype == 1, "pack_views: dtype must be 0 (bf16) or 1 (fp16)");
  MMLF_REQUIRE(views && out, "pack_views: null buffer");
  MMLF_REQUIRE(ld % 8 == 0 && ld >= C, "pack_views: ld %d must be a multiple of 8 and >= C %d", ld, C);
  MMLF_REQUIRE(static_cast<int64_t>(B) * (H + 1) <= 65535 * 1024ll, "pack_views: batch too large");
  PackParams p;
  p.views = views; p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.B = B; p.C = C; p.H = H; p.W = W; p.ld = ld;
  p.do_shift = do_shift; p.stack = stack; p.n = n; p.dtype = dtype;
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
    dim3 grid(ceil_div(W + 1, 32), static_cast<unsigned>(nb * (H + 1)));
    pack_views_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(q);
    if (int rc = check_launch("pack_views_kernel")) return rc;
  }
  return 0;
}

extern "C" int mmlf_pack_views(const float* views, int B, int C, int H, int W, void* out, int ld, int dtype,
                               void* stream) {
  return launch_pack(views, B, C, H, W, out, ld, dtype, 0, 0, 0, 0.0, stream);
}

extern "C" int mmlf_shift_pack(const float* src, int stack, int B, int n, int H, int W, double disp, void* out,
                               int ld, int dtype, void* stream) {
  MMLF_REQUIRE(stack >= 0 && stack < 4, "shift_pack: stack must be 0..3");
  MMLF_REQUIRE(n >= 1 && n <= 16, "shift_pack: n must be in [1, 16]");
  return launch_pack(src, B, n * 3, H, W, out, ld, dtype, 1, stack, n, disp, stream);
}
