// Training augmentation chain on the GPU: the reference composes, per sample and on CPU workers,
//   RandomDownSampling -> RandomShift -> RandomCrop(ps + 16) -> CenterCrop(ps) -> RandomRotate -> RedistColor ->
//   Brightness -> Contrast        (/root/reference/mmlf/train/cli.py:78-87, /root/reference/mmlf/data/hci4d.py)
// and ships 2 GB of float32 patches per 512-patch step to the GPU.  Here the scenes stay resident in HBM and one gather
// kernel evaluates the whole chain per output pixel from explicit per-sample parameters (drawn on the host in the
// reference's order), with the reference's arithmetic (NumPy >= 2 promotion: float32 lerps and scalings, float64
// colour products rounded to float32 after every accumulation) -- bit exact up to Contrast's mean, which NumPy sums
// pairwise in float32 and this file sums in float64.
#include <math.h>

#include "../../include/mmlf_b200.h"
#include "common.cuh"
#include "host_util.h"

static_assert(sizeof(mmlf_aug_sample) == 408, "mmlf_aug_sample layout (mirrored by mmlf_b200/data/augment.py::AugSample)");

namespace mmlf {

__device__ __forceinline__ int aug_src_index(int j, int s, int n, int sign) {      // as Shift: python slice semantics
  if (s == 0 || s >= n || -s >= n) return j;
  int r = j - sign * s;
  if (r < 0) r += n;
  if (r >= n) r -= n;
  return r;
}
__device__ __forceinline__ float aug_lerp2(float a, float w0, float b, float w1) {
  return __fadd_rn(__fmul_rn(a, w0), __fmul_rn(b, w1));
}
// patch coordinates before `r` rotations: out[y][x] = in[T^r(y, x)], T(y, x) = (x, n - 1 - y)   (hci4d.py:1057-1060)
__device__ __forceinline__ void aug_unrotate(int r, int n, int y, int x, int& Y, int& X) {
  Y = y; X = x;
  for (int i = 0; i < r; ++i) {
    const int t = Y;
    Y = X;
    X = n - 1 - t;
  }
}

// One block = a 16 x 16 output tile of ONE view (3 colours) of one output stack, or of the centre view (stack index 4):
// stack, view, taps and every sample parameter are block-uniform (the first version decoded stack / view / pixel from a
// flat index with three run-time divisions per thread).  Threads are laid out along the SOURCE row: for an odd number of
// quarter turns the source column runs along the output y, so the tile is computed transposed and turned in shared
// memory -- loads and stores both touch whole 64-byte runs (the flat version stored 19 sectors per request on average).
constexpr int kAugTile = 16;
__global__ void __launch_bounds__(256)
augment_views_kernel(const float* __restrict__ stacks, const float* __restrict__ center, int n, int H, int W,
                     const mmlf_aug_sample* __restrict__ samples, int B, int ps, float* __restrict__ out_views,
                     float* __restrict__ out_center, double* __restrict__ view_sums) {
  __shared__ float turn[3][kAugTile][kAugTile + 1];
  const int b = blockIdx.y;
  const mmlf_aug_sample& sp = samples[b];
  const int tiles = (ps + kAugTile - 1) / kAugTile;
  const int plane_id = blockIdx.x / (tiles * tiles);      // 0 .. 4n - 1: (stack, view); 4n: the centre view
  const int tile_id = blockIdx.x - plane_id * tiles * tiles;
  const int so = plane_id < 4 * n ? plane_id / n : 4;
  const int k = so < 4 ? plane_id - so * n : 0;
  const int ty0 = (tile_id / tiles) * kAugTile, tx0 = (tile_id % tiles) * kAugTile;
  const int tq = threadIdx.x >> 4, tt = threadIdx.x & 15;  // tt runs along the source row
  const bool odd = sp.r & 1;
  const int y = ty0 + (odd ? tt : tq), x = tx0 + (odd ? tq : tt);
  const bool inside = y < ps && x < ps;
  double hsum = 0.0;
  float o[3] = {0.f, 0.f, 0.f};
  if (inside) {
    int Y, X;
    aug_unrotate(sp.r, ps, y, x, Y, X);
    const int f = sp.f;
    const int Hd = (H + f - 1) / f, Wd = (W + f - 1) / f;
    const int yy = sp.cy + Y, xx = sp.cx + X;             // in the down-sampled (and shifted) image
    const int64_t plane = static_cast<int64_t>(H) * W;
    float v[3];
    if (so == 4) {
      const float* src = center + static_cast<int64_t>(sp.scene) * 3 * plane;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = __ldg(src + c * plane + static_cast<int64_t>(yy * f) * W + xx * f);
    } else {
      const int si = sp.src[so];
      const int ki = sp.flip[so] ? n - 1 - k : k;
      const float* src = stacks + ((static_cast<int64_t>(sp.scene) * 4 + si) * n + ki) * 3 * plane;
      const float w0 = sp.w0[ki], w1 = sp.w1[ki];
      const int s0 = sp.s0[ki], s1 = sp.s1[ki];
      const bool has_w = si != 1, has_v = si != 0;
      const int vsign = si == 2 ? -1 : +1;
      const int r0 = has_v ? aug_src_index(yy, s0, Hd, vsign) : yy;
      const int r1 = has_v ? aug_src_index(yy, s1, Hd, vsign) : yy;
      const int c0 = has_w ? aug_src_index(xx, s0, Wd, +1) : xx;
      const int c1 = has_w ? aug_src_index(xx, s1, Wd, +1) : xx;
      // all loads of the pixel first (element offsets inside a plane fit 32 bits)
      const int oa0 = r0 * f * W + c0 * f, oa1 = r0 * f * W + c1 * f, ob0 = r1 * f * W + c0 * f, ob1 = r1 * f * W + c1 * f;
      float a0[3], a1[3], b0[3], b1[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* pl = src + c * plane;
        a0[c] = __ldg(pl + oa0);
        a1[c] = has_w ? __ldg(pl + oa1) : 0.f;
        b0[c] = has_v ? __ldg(pl + ob0) : 0.f;
        b1[c] = has_v && has_w ? __ldg(pl + ob1) : 0.f;
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float t0 = has_w ? aug_lerp2(a0[c], w0, a1[c], w1) : a0[c];
        if (!has_v) {
          v[c] = t0;
        } else {
          const float t1 = has_w ? aug_lerp2(b0[c], w0, b1[c], w1) : b0[c];
          v[c] = aug_lerp2(t0, w0, t1, w1);
        }
      }
    }
    // RedistColor (float64 products, float32 after every accumulation), then Brightness (float32)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float acc = static_cast<float>(__dmul_rn(sp.mat[3 * c + 0], static_cast<double>(v[0])));
      acc = static_cast<float>(__dadd_rn(static_cast<double>(acc), __dmul_rn(sp.mat[3 * c + 1], static_cast<double>(v[1]))));
      acc = static_cast<float>(__dadd_rn(static_cast<double>(acc), __dmul_rn(sp.mat[3 * c + 2], static_cast<double>(v[2]))));
      o[c] = __fmul_rn(acc, sp.bright);
    }
    if (so == 0) hsum = static_cast<double>(o[0]) + static_cast<double>(o[1]) + static_cast<double>(o[2]);
  }
  // store: thread (tq, tt) writes output pixel (ty0 + tq, tx0 + tt); transposed tiles go through shared memory
  int oy = y, ox = x;
  if (odd) {                                              // block-uniform
#pragma unroll
    for (int c = 0; c < 3; ++c) turn[c][tt][tq] = o[c];   // [output y - ty0][output x - tx0]
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = turn[c][tq][tt];
    oy = ty0 + tq;
    ox = tx0 + tt;
  }
  if (oy < ps && ox < ps) {
    const int64_t pp = static_cast<int64_t>(ps) * ps;
    float* dst = so == 4 ? out_center + static_cast<int64_t>(b) * 3 * pp
                         : out_views + (((static_cast<int64_t>(so) * B + b) * n + k) * 3) * pp;
    dst += oy * ps + ox;
#pragma unroll
    for (int c = 0; c < 3; ++c) dst[c * pp] = o[c];
  }
  // Contrast's mean is over the h stack (data[0]) after Brightness: block partial -> one atomic
  if (so != 0) return;                                    // block-uniform
  __shared__ double red[8];
  hsum = warp_sum(hsum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = hsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += red[i];
    if (t != 0.0) atomicAdd(view_sums + b, t);
  }
}

// x * contrast + mean * (1 - contrast) over the four stacks and the centre view of a sample.  The mean (a float64
// division) is derived once per block -- per thread it made this elementwise pass SFU bound (418 us for 64 patches) -- and
// a thread handles four consecutive floats (V = 4) when the plane size allows 128-bit accesses.
template <int V>
__global__ void __launch_bounds__(256)
augment_contrast_kernel(float* __restrict__ views, float* __restrict__ center, const mmlf_aug_sample* __restrict__ samples,
                        const double* __restrict__ view_sums, const float* __restrict__ mean_override, int B, int n,
                        int ps) {
  const int b = blockIdx.y;
  const int per_stack = n * 3 * ps * ps;
  const int total = 4 * per_stack + 3 * ps * ps;
  __shared__ float s_off;
  if (threadIdx.x == 0) {
    const float mean = mean_override ? mean_override[b] : static_cast<float>(view_sums[b] / static_cast<double>(per_stack));
    s_off = __fmul_rn(mean, samples[b].one_minus_contrast);
  }
  __syncthreads();
  const int idx = (blockIdx.x * blockDim.x + threadIdx.x) * V;
  if (idx >= total) return;
  const float a = samples[b].contrast;
  const float off = s_off;
  float* p;
  if (idx < 4 * per_stack) {
    const int so = idx / per_stack, rem = idx - so * per_stack;
    p = views + (static_cast<int64_t>(so) * B + b) * per_stack + rem;
  } else {
    p = center + static_cast<int64_t>(b) * 3 * ps * ps + (idx - 4 * per_stack);
  }
  if (V == 4) {
    float4 x = *reinterpret_cast<float4*>(p);
    x.x = __fadd_rn(__fmul_rn(x.x, a), off); x.y = __fadd_rn(__fmul_rn(x.y, a), off);
    x.z = __fadd_rn(__fmul_rn(x.z, a), off); x.w = __fadd_rn(__fmul_rn(x.w, a), off);
    *reinterpret_cast<float4*>(p) = x;
  } else {
    *p = __fadd_rn(__fmul_rn(*p, a), off);
  }
}

// gt, mpi (rotated like the views) and mask (cropped only)
__global__ void __launch_bounds__(256)
augment_targets_kernel(const float* __restrict__ gt, const float* __restrict__ mpi, const int32_t* __restrict__ mask, int K,
                       int H, int W, const mmlf_aug_sample* __restrict__ samples, int ps, float* __restrict__ out_gt,
                       float* __restrict__ out_mpi, int32_t* __restrict__ out_mask) {
  const int b = blockIdx.y;
  const mmlf_aug_sample& sp = samples[b];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ps * ps) return;
  const int y = idx / ps, x = idx - y * ps;
  int Y, X;
  aug_unrotate(sp.r, ps, y, x, Y, X);
  const int f = sp.f;
  const int64_t plane = static_cast<int64_t>(H) * W;
  const int64_t src = static_cast<int64_t>((sp.cy + Y) * f) * W + (sp.cx + X) * f;
  const int64_t pp = static_cast<int64_t>(ps) * ps;
  if (out_gt) {
    // gt /= f; gt -= disp  (float32 array, python floats: hci4d.py:506, 984-985)
    const float g = __fdiv_rn(__ldg(gt + sp.scene * plane + src), static_cast<float>(f));
    out_gt[b * pp + idx] = __fsub_rn(g, sp.disp_f);
  }
  if (out_mpi) {
    for (int k = 0; k < K; ++k) {
      const float* m = mpi + (static_cast<int64_t>(sp.scene) * K + k) * 5 * plane + src;
      float* o = out_mpi + (static_cast<int64_t>(b) * K + k) * 5 * pp + idx;
      for (int c = 0; c < 4; ++c) o[c * pp] = __ldg(m + c * plane);
      // float64 array in the reference: (d / f) - disp in double, rounded once by the later .float()
      const double d = __dsub_rn(__ddiv_rn(static_cast<double>(__ldg(m + 4 * plane)), static_cast<double>(f)), sp.disp);
      o[4 * pp] = static_cast<float>(d);
    }
  }
  if (out_mask) {
    const int64_t ms = static_cast<int64_t>((sp.cy + y) * f) * W + (sp.cx + x) * f;      // the mask is NOT rotated
    out_mask[b * pp + idx] = __ldg(mask + sp.scene * plane + ms);
  }
}

}  // namespace mmlf

using namespace mmlf;

extern "C" int mmlf_augment_fill(mmlf_aug_sample* s, int r, double disp, int n) {
  MMLF_REQUIRE(s != nullptr && r >= 0 && r <= 3 && n >= 1 && n <= 16, "augment_fill: bad arguments");
  s->r = r;
  int src[4] = {0, 1, 2, 3}, flip[4] = {0, 0, 0, 0};
  for (int i = 0; i < r; ++i) {                      // Rotate90: h' = v, v' = flip(h), i' = d, d' = flip(i)
    const int ns[4] = {src[1], src[0], src[3], src[2]};
    const int nf[4] = {flip[1], flip[0] ^ 1, flip[3], flip[2] ^ 1};
    for (int k = 0; k < 4; ++k) { src[k] = ns[k]; flip[k] = nf[k]; }
  }
  for (int k = 0; k < 4; ++k) { s->src[k] = src[k]; s->flip[k] = flip[k]; }
  s->disp = disp;
  s->disp_f = static_cast<float>(disp);
  return mmlf_shift_taps(disp, n, s->w0, s->w1, s->s0, s->s1);
}

extern "C" int mmlf_augment_patches(const float* stacks, const float* center, const float* gt, const float* mpi,
                                    const int32_t* mask, int S, int n, int K, int H, int W, const mmlf_aug_sample* samples,
                                    int B, int ps, float* out_views, float* out_center, double* view_sums, float* out_gt,
                                    float* out_mpi, int32_t* out_mask, void* stream) {
  MMLF_REQUIRE(stacks && center && samples && out_views && out_center && view_sums, "augment_patches: null buffer");
  MMLF_REQUIRE(S >= 1 && n >= 1 && n <= 16 && B >= 1 && B <= 65535 && ps >= 1, "augment_patches: bad sizes");
  MMLF_REQUIRE((!out_gt || gt) && (!out_mpi || (mpi && K >= 1)) && (!out_mask || mask), "augment_patches: missing source");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int tiles = ceil_div(ps, 16);
  const int blocks = (4 * n + 1) * tiles * tiles;               // 16 x 16 output tiles, one view plane per group of blocks
  augment_views_kernel<<<dim3(blocks, B), 256, 0, st>>>(stacks, center, n, H, W, samples, B, ps, out_views,
                                                                      out_center, view_sums);
  if (int rc = check_launch("augment_views_kernel")) return rc;
  if (out_gt || out_mpi || out_mask) {
    augment_targets_kernel<<<dim3(ceil_div(ps * ps, 256), B), 256, 0, st>>>(gt, mpi, mask, K, H, W, samples, ps, out_gt,
                                                                            out_mpi, out_mask);
    return check_launch("augment_targets_kernel");
  }
  return 0;
}

extern "C" int mmlf_augment_contrast(float* views, float* center, const mmlf_aug_sample* samples, const double* view_sums,
                                     const float* mean_override, int B, int n, int ps, void* stream) {
  MMLF_REQUIRE(views && center && samples && (view_sums || mean_override), "augment_contrast: null buffer");
  const int total = 4 * n * 3 * ps * ps + 3 * ps * ps;
  // 128-bit path: every stack / centre block starts on a multiple of 4 floats and the buffers are 16-byte aligned
  const bool vec = (ps * ps) % 4 == 0 && reinterpret_cast<uintptr_t>(views) % 16 == 0 &&
                   reinterpret_cast<uintptr_t>(center) % 16 == 0;
  if (vec)
    augment_contrast_kernel<4><<<dim3(ceil_div(total / 4, 256), B), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        views, center, samples, view_sums, mean_override, B, n, ps);
  else
    augment_contrast_kernel<1><<<dim3(ceil_div(total, 256), B), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        views, center, samples, view_sums, mean_override, B, n, ps);
  return check_launch("augment_contrast_kernel");
}
