// 2x2 convolution as an implicit GEMM on the 5th-gen tensor cores (tcgen05 + TMEM), fed by TMA.
//
// Replaces nn.Conv2d(cin, cout, 2, padding=1|0) of /root/reference/mmlf/model/feed_forward.py:123,125
// (forward) and, with dgrad-packed weights, its data gradient.  See DESIGN.md section 4.
//
//   GEMM view : D[slot][n] = sum_{tap, c} A[slot + off(tap)][c] * Wp[n][tap][c]
//   A operand : activations in the 16-bit slot layout; one (tap, 64-channel chunk) = one 2-D TMA box of
//               128 rows x 128 B landing as a canonical K-major SWIZZLE_128B tile.  Rows outside the array and
//               channels >= cin_pad are zero-filled by TMA (that is the conv padding of the first/last image).
//   B operand : packed weights [n_pad][4 * kc * 64], K-major, SWIZZLE_128B, re-streamed from L2 per tile.
//   D         : fp32 in TMEM, 2 x 128 lanes x n_pad columns per CTA pair, <= 2 MMAs of N <= 256 per k-step.
//   CTA pair  : cta_group::2, M = 256 slots per tile; each CTA loads its own 128 A rows and HALF of the weight rows
//               of every MMA (the hardware reads the other half from the peer's shared memory).
//   roles     : warp 0 = TMA producer, warp 1 = MMA issuer (leader CTA) + TMEM owner, warps 2..5 = epilogue (one
//               TMEM lane quadrant = 32 slots each).  Persistent, static round-robin over 256-slot tiles.
//   epilogue  : TMEM -> registers -> (+bias | *scale+shift, ReLU, ReLU-bit gate, halo mask) -> 16-bit ->
//               per-warp swizzled shared-memory staging -> TMA store (32 slots x 32 channels per box); optional
//               second copy in the other 16-bit format, ReLU sign bits, per-channel sum / sum of squares.
#include "../../include/mmlf_b200.h"
#include "common.cuh"
#include <stdlib.h>

#include "host_util.h"

namespace mmlf {

constexpr int kTileM = 128;
constexpr int kABytes = kTileM * 128;  // one A stage: 128 rows x 64 16-bit channels
constexpr int kConvThreads = 192;
constexpr int kMaxStages = 8;
constexpr uint32_t kTmemCols = 512;
constexpr int kBoxBytes = 32 * 64;     // epilogue staging box: 32 slots x 32 channels x 2 B (SWIZZLE_64B)
constexpr int kBitsPitch = 11;         // words per slot row of the per-warp bit scratch (odd: conflict-free)
constexpr int kMaxNPad = 320;

struct ConvParams {
  int64_t n_slots;
  int Hp, Wp, H, W;
  int n_pad, n_parts, n_part;
  int n_kc, last_ksteps;
  int tap_off[4];
  int num_tiles, stages;
  int type, relu, out_mode, n_real, ld_out;
  int ab_dtype, out_dtype, out2_dtype;
  int has_scale, dual, ld_bits;
  uint32_t epi_off, bits_off, stat_off, aux_off;   // shared-memory offsets from the 1024-aligned base
  const float* bias;
  const float* scale;
  const float* shift;
  const uint32_t* gate_bits;
  uint32_t* relu_bits;
  double* col_sums;
  void* out;
  long long* stats;                             // optional per-CTA cycle counters (debug): [grid][8]
};

__device__ __forceinline__ void slot_coords(const ConvParams& p, int64_t s, int& b, int& sy, int& sx) {
  const int per_img = p.Hp * p.Wp;
  b = static_cast<int>(s / per_img);
  const int rem = static_cast<int>(s - static_cast<int64_t>(b) * per_img);
  sy = rem / p.Wp;
  sx = rem - sy * p.Wp;
}

// Epilogue math for `n` (<= 32) consecutive output channels [c0, c0+n) of one slot (CUDA-core reference kernel and the
// direct-store modes of the tensor-core kernel share it).
__device__ __forceinline__ float epi_value(const ConvParams& p, float acc, float add, float mul, bool valid) {
  float x = p.has_scale ? fmaf(acc, mul, add) : acc + add;
  if (p.relu) x = fmaxf(x, 0.f);
  return valid ? x : 0.f;
}

// Direct (non-TMA) stores of one slot's channels [c0, c0 + 16): fp32 slot rows (out_mode 1) or planar fp32 (out_mode 2)
__device__ __forceinline__ void direct_store16(const ConvParams& p, int64_t s, bool in_range, bool valid, int b, int sy,
                                               int sx, int c0, const float* v) {
  if (!in_range) return;
  if (p.out_mode == 1) {
    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + s * p.ld_out + c0);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else {
    if (!valid) return;
    const int Ho = p.type ? p.H : p.Hp, Wo = p.type ? p.W : p.Wp;
    const int y = p.type ? sy - 1 : sy, x = p.type ? sx - 1 : sx;
    float* o = reinterpret_cast<float*>(p.out) + ((static_cast<int64_t>(b) * p.n_real + c0) * Ho + y) * Wo + x;
    const int64_t plane = static_cast<int64_t>(Ho) * Wo;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c0 + j < p.n_real) o[j * plane] = v[j];
  }
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
      "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 16-column variant into the lower half of a 32-register box (upper half zeroed)
__device__ __forceinline__ void tmem_ld16_lo(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int j = 16; j < 32; ++j) r[j] = 0u;
}

__device__ __forceinline__ uint32_t pack_f16x2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// One epilogue warp's context for the staged (out_mode 0) path
struct EpiWarp {
  uint32_t stg_addr;        // shared address of this warp's staging boxes: [2 (ring)][1 + dual] x kBoxBytes
  uint32_t* s_gate;         // [32][kBitsPitch] gate bits of the current tile (or nullptr)
  uint32_t* s_relu;         // [32][kBitsPitch] relu bits of the current tile (or nullptr)
  float* s_stat;            // [2][n_pad] per-warp partial column sums (or nullptr)
  const float* s_add;
  const float* s_mul;
  uint32_t ring;            // boxes stored so far (selects the staging buffer)
};

// Processes one box of 32 output channels [box * 32, box * 32 + 32) of the warp's 32 slots: math, 16-bit packing,
// swizzled staging, TMA store, statistics.  r holds the fp32 accumulators of this thread's slot.
__device__ __forceinline__ void epi_box_staged(const ConvParams& p, EpiWarp& w, const CUtensorMap* tmap_o,
                                               const CUtensorMap* tmap_o2, const uint32_t (&r)[32], int box,
                                               int row0, bool valid, int lane) {
  const int c0 = box * 32;
  float x[32];
  {
    const float4* add4 = reinterpret_cast<const float4*>(w.s_add + c0);
    if (p.has_scale) {
      const float4* mul4 = reinterpret_cast<const float4*>(w.s_mul + c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 a = add4[j], m = mul4[j];
        x[4 * j + 0] = fmaf(__uint_as_float(r[4 * j + 0]), m.x, a.x);
        x[4 * j + 1] = fmaf(__uint_as_float(r[4 * j + 1]), m.y, a.y);
        x[4 * j + 2] = fmaf(__uint_as_float(r[4 * j + 2]), m.z, a.z);
        x[4 * j + 3] = fmaf(__uint_as_float(r[4 * j + 3]), m.w, a.w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 a = add4[j];
        x[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + a.x;
        x[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + a.y;
        x[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + a.z;
        x[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + a.w;
      }
    }
  }
  if (p.relu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
  }
  if (w.s_gate) {
    const uint32_t g = w.s_gate[lane * kBitsPitch + box];
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = ((g >> j) & 1u) ? x[j] : 0.f;
  }
  if (!valid) {
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = 0.f;
  }
  if (w.s_relu) {
    uint32_t bits = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) bits |= (x[j] > 0.f ? 1u : 0u) << j;
    w.s_relu[lane * kBitsPitch + box] = bits;
  }

  // staging buffer of this box; the TMA store that used it two boxes ago must have finished reading it
  const uint32_t buf = w.stg_addr + (w.ring & 1u) * (p.dual ? 2u * kBoxBytes : kBoxBytes);
  if (lane == 0) bulk_wait_read<1>();
  __syncwarp();
  // slot row `lane` occupies 64 B; SWIZZLE_64B: 16-byte chunk j lands at chunk j ^ ((row >> 1) & 3)
  const uint32_t row_addr = buf + lane * 64;
  const uint32_t sw = (lane >> 1) & 3;
  {
    uint32_t pk[16];
    if (p.out_dtype == kFP16) {
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = pack_f16x2_sat(x[2 * j], x[2 * j + 1]);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(x[2 * j], x[2 * j + 1]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + ((j ^ sw) << 4)), "r"(pk[4 * j]),
                   "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                   : "memory");
  }
  if (p.dual) {
    uint32_t pk[16];
    if (p.out2_dtype == kFP16) {
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = pack_f16x2_sat(x[2 * j], x[2 * j + 1]);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(x[2 * j], x[2 * j + 1]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + kBoxBytes + ((j ^ sw) << 4)),
                   "r"(pk[4 * j]), "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                   : "memory");
  }
  fence_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_2d(tmap_o, buf, c0, row0);
    if (p.dual) tma_store_2d(tmap_o2, buf + kBoxBytes, c0, row0);
    bulk_commit();
  }
  ++w.ring;

  if (w.s_stat) {
    // per-channel sum / sum of squares of the stored (rounded) values: lane handles channel pair (lane & 15) of the
    // 16 slots with parity (lane >> 4); rows 2k and 2k+1 share a 128-byte line, so the two half-warps never conflict
    const int cp = lane & 15, par = lane >> 4;
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int row = 2 * k + par;
      const uint32_t a = buf + row * 64 + ((((cp >> 2) ^ ((row >> 1) & 3))) << 4) + ((cp & 3) << 2);
      uint32_t v;
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
      float lo, hi;
      unpack16x2(v, p.out_dtype, lo, hi);
      s0 += lo;
      s1 += hi;
      q0 = fmaf(lo, lo, q0);
      q1 = fmaf(hi, hi, q1);
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 16);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    q0 += __shfl_xor_sync(0xffffffffu, q0, 16);
    q1 += __shfl_xor_sync(0xffffffffu, q1, 16);
    float2* dst = reinterpret_cast<float2*>(w.s_stat + (par ? p.n_pad : 0) + c0 + 2 * cp);
    if (c0 + 2 * cp < p.n_pad) {
      float2 cur = *dst;
      cur.x += par ? q0 : s0;
      cur.y += par ? q1 : s1;
      *dst = cur;
    }
  }
}

// Direct-store variant (fp32 slot rows / planar fp32 outputs of the heads)
__device__ __forceinline__ void epi_box_direct(const ConvParams& p, const float* s_add, const float* s_mul,
                                               const uint32_t (&r)[32], int box, int ncols, int64_t s, bool in_range,
                                               bool valid, int b, int sy, int sx) {
  const int c0 = box * 32;
  float x[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) x[j] = epi_value(p, __uint_as_float(r[j]), s_add[c0 + j], s_mul[c0 + j], valid);
  direct_store16(p, s, in_range, valid, b, sy, sx, c0, x);
  if (ncols > 16) direct_store16(p, s, in_range, valid, b, sy, sx, c0 + 16, x + 16);
}

// ---------------------------------------------------------------------------------------------------------------
// TMEM plan: consecutive tiles alternate between two accumulator regions so that the epilogue of tile t overlaps the
// MMAs of tile t+1.  Two 288-column accumulators do not fit into 512 columns, so the regions are [0, n_pad) and
// [512 - n_pad, 512) and share `ovl` = 2 * n_pad - 512 columns; the epilogue pulls the shared columns into registers
// first and releases them through a separate barrier, after which the next tile's MMAs may start.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
conv2x2_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                   const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ CUtensorMap tmap_o2,
                   const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t tiles_addr = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (tiles_addr - raw_addr);
  const uint32_t b_bytes = static_cast<uint32_t>(p.n_pad) * 64u;          // half of the weight rows per CTA
  const uint32_t stage_bytes = kABytes + b_bytes;

  uint8_t* aux = smem + p.aux_off;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full_bar = empty_bar + kMaxStages;      // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;          // [2], leader's copy is used
  uint64_t* tmem_ovl_bar = tmem_empty_bar + 2;           // leader's copy is used
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_ovl_bar + 1);
  float* s_add = reinterpret_cast<float*>(aux + 192);   // 16-byte aligned (read as float4)
  float* s_mul = s_add + kMaxNPad;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const bool leader = rank == 0;
  // the overlap trick needs the shared columns to be whole 32-channel boxes; otherwise the regions are used strictly
  // one after the other (ovl_wait_all)
  const int ovl_cols = p.n_pad > 256 ? 2 * p.n_pad - 512 : 0;
  const bool ovl_ok = (ovl_cols & 31) == 0 && (p.n_pad & 31) == 0;
  const int base1 = p.n_pad > 256 ? 512 - p.n_pad : 256;      // TMEM column base of odd tiles
  const int ovl = ovl_ok ? ovl_cols : 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    if (p.out_mode == 0) prefetch_tmap(&tmap_o);
    if (p.dual) prefetch_tmap(&tmap_o2);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < p.stages; ++i) {
        mbar_init(smem_u32(&full_bar[i]), 1);          // leader: one arrive.expect_tx covering both CTAs' bytes
        mbar_init(smem_u32(&empty_bar[i]), 1);         // one multicast commit from the leader's MMA thread
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(smem_u32(&tmem_full_bar[i]), 1);
        mbar_init(smem_u32(&tmem_empty_bar[i]), 8);    // one elected lane of the 4 epilogue warps of both CTAs
      }
      mbar_init(smem_u32(tmem_ovl_bar), 8);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_pair(smem_u32(tmem_ptr_smem), kTmemCols);
  }
  for (int i = threadIdx.x; i < kMaxNPad; i += kConvThreads) {
    float add = 0.f, mul = 1.f;
    if (i < p.n_pad) {
      if (p.has_scale) {
        mul = p.scale[i];
        add = p.shift[i] + (p.bias ? p.bias[i] * mul : 0.f);
      } else if (p.bias) {
        add = p.bias[i];
      }
    }
    s_add[i] = add;
    s_mul[i] = mul;
  }
  if (p.stat_off) {
    float* st = reinterpret_cast<float*>(smem + p.stat_off);
    for (int i = threadIdx.x; i < 4 * 2 * p.n_pad; i += kConvThreads) st[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();                                      // peer barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int half_rows = p.n_part >> 1;                 // weight rows each CTA supplies per MMA

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer.  The whole warp runs the loop so
    // that every operand is warp-uniform (UTMALDG takes uniform registers; a lane-0-only region makes the compiler
    // wrap each instruction in an elect/broadcast loop); one elected lane issues.
    uint32_t stage = 0, phase = 0;
    long long t_wait = 0, t_begin = clock64();
    for (int tile = pair; tile < p.num_tiles; tile += n_pairs) {
      const int row0 = tile * (2 * kTileM) + static_cast<int>(rank) * kTileM;
      for (int tap = 0; tap < 4; ++tap) {
        for (int kc = 0; kc < p.n_kc; ++kc) {
          const long long tw = clock64();
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
          t_wait += clock64() - tw;
          const uint32_t fb = smem_u32(&full_bar[stage]);
          const uint32_t a_dst = tiles_addr + stage * stage_bytes;
          const uint32_t b_dst = a_dst + kABytes;
          const int kcol = (tap * p.n_kc + kc) * 64;
          if (elect_one()) {
            if (leader) mbar_arrive_expect_tx(fb, 2 * stage_bytes);
            tma_load_2d_pair(a_dst, &tmap_a, fb, kc * 64, row0 + p.tap_off[tap], kEvictNormal);
            for (int part = 0; part < p.n_parts; ++part)
              tma_load_2d_pair(b_dst + part * half_rows * 128, &tmap_b, fb, kcol,
                               part * p.n_part + static_cast<int>(rank) * half_rows, kEvictLast);
          }
          __syncwarp();
          if (++stage == static_cast<uint32_t>(p.stages)) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    if (p.stats && lane == 0) {
      p.stats[blockIdx.x * 8 + 0] = clock64() - t_begin;   // producer total
      p.stats[blockIdx.x * 8 + 1] = t_wait;                // producer waiting for free stages
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA; warp-uniform loop,
    // one elected lane issues the MMAs and the commits that track them)
    if (leader) {
      const uint32_t idesc = make_idesc_16(2 * kTileM, p.n_part, 0, 0, p.ab_dtype, p.ab_dtype);
      const uint32_t part_bytes = static_cast<uint32_t>(half_rows) * 128u;
      uint32_t stage = 0, phase = 0;
      int it = 0;
      long long t_full = 0, t_tmem = 0, t_begin = clock64();
      for (int tile = pair; tile < p.num_tiles; tile += n_pairs, ++it) {
        const int par = it & 1;
        long long tw = clock64();
        mbar_wait(smem_u32(&tmem_empty_bar[par]), ((it >> 1) & 1) ^ 1u);   // region drained (tile it - 2)
        // regions overlap only when n_pad > 256: then the columns shared with tile it - 1 must have been drained
        if (p.n_pad > 256 && it > 0) mbar_wait(smem_u32(tmem_ovl_bar), (it - 1) & 1);
        t_tmem += clock64() - tw;
        tc_fence_after();
        const uint32_t acc_base = tmem_base + (par ? base1 : 0);
        uint32_t accumulate = 0;
        for (int tap = 0; tap < 4; ++tap) {
          for (int kc = 0; kc < p.n_kc; ++kc) {
            tw = clock64();
            mbar_wait(smem_u32(&full_bar[stage]), phase);
            t_full += clock64() - tw;
            tc_fence_after();
            const uint32_t a_addr = tiles_addr + stage * stage_bytes;
            const uint32_t b_addr = a_addr + kABytes;
            const int ksteps = (kc == p.n_kc - 1) ? p.last_ksteps : 4;
            const uint64_t adesc0 = make_sw128_desc(a_addr, 0, 1024);
            const uint64_t bdesc0 = make_sw128_desc(b_addr, 0, 1024);
            const bool last = (tap == 3 && kc == p.n_kc - 1);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (k < ksteps) {
                  // advancing by k * 32 bytes inside the 128-byte swizzle row = +2k in the (addr >> 4) field
                  umma_f16_pair(acc_base, adesc0 + 2 * k, bdesc0 + 2 * k, idesc, accumulate);
                  if (p.n_parts == 2)
                    umma_f16_pair(acc_base + p.n_part, adesc0 + 2 * k, bdesc0 + (part_bytes >> 4) + 2 * k, idesc,
                                  accumulate);
                  accumulate = 1;
                }
              }
              umma_commit_pair(smem_u32(&empty_bar[stage]));
              if (last) umma_commit_pair(smem_u32(&tmem_full_bar[par]));
            }
            accumulate = 1;
            __syncwarp();
            if (++stage == static_cast<uint32_t>(p.stages)) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
      if (p.stats && lane == 0) {
        p.stats[blockIdx.x * 8 + 2] = clock64() - t_begin;   // MMA issuer total
        p.stats[blockIdx.x * 8 + 3] = t_full;                // ... waiting for TMA data
        p.stats[blockIdx.x * 8 + 4] = t_tmem;                // ... waiting for the epilogue to free TMEM
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: one warp per TMEM lane quadrant
    const int q = warp & 3;
    const bool staged = p.out_mode == 0;
    EpiWarp w;
    w.stg_addr = tiles_addr + p.epi_off + q * (p.dual ? 4u : 2u) * kBoxBytes;
    uint32_t* bits_base = reinterpret_cast<uint32_t*>(smem + p.bits_off) + q * 2 * 32 * kBitsPitch;
    w.s_gate = p.gate_bits ? bits_base : nullptr;
    w.s_relu = p.relu_bits ? bits_base + 32 * kBitsPitch : nullptr;
    w.s_stat = p.stat_off ? reinterpret_cast<float*>(smem + p.stat_off) + q * 2 * p.n_pad : nullptr;
    w.s_add = s_add;
    w.s_mul = s_mul;
    w.ring = 0;
    const int n_boxes = (p.n_pad + 31) >> 5;
    const int last_cols = p.n_pad - (n_boxes - 1) * 32;        // 32 or 16
    const int ovl_boxes = ovl >> 5;
    const int bits_words = 32 * p.ld_bits;                     // words of bit rows per warp and tile
    int it = 0;
    long long t_wait = 0, t_begin = clock64();
    for (int tile = pair; tile < p.num_tiles; tile += n_pairs, ++it) {
      const int par = it & 1;
      const int row0 = tile * (2 * kTileM) + static_cast<int>(rank) * kTileM + q * 32;
      const int64_t s = static_cast<int64_t>(row0) + lane;
      const bool in_range = s < p.n_slots;
      int b = 0, sy = 0, sx = 0;
      if (in_range) slot_coords(p, s, b, sy, sx);
      const bool valid = in_range && (p.type == 0 || (sy >= 1 && sx >= 1));
      if (w.s_gate) {
        // the warp's 32 gate rows are contiguous in global memory: coalesced copy into the scratch
        const uint32_t* g = p.gate_bits + static_cast<int64_t>(row0) * p.ld_bits;
        const int64_t lim = (p.n_slots - row0) * p.ld_bits;
        for (int i = lane; i < bits_words; i += 32) {
          const int rr = i / p.ld_bits, ww = i - rr * p.ld_bits;
          w.s_gate[rr * kBitsPitch + ww] = i < lim ? __ldg(g + i) : 0u;
        }
        __syncwarp();
      }
      const long long tw = clock64();
      mbar_wait(smem_u32(&tmem_full_bar[par]), (it >> 1) & 1);
      t_wait += clock64() - tw;
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (par ? base1 : 0);
      // box order: the columns shared with the other region first (the last boxes of an even tile, the first ones of
      // an odd tile) -- a rotation of the natural order
      const int first = (par == 0 && ovl_boxes > 0) ? n_boxes - ovl_boxes : 0;
      auto box_at = [&](int i) { int bx = first + i; return bx >= n_boxes ? bx - n_boxes : bx; };
      auto load_box = [&](int bx, uint32_t (&r)[32]) {
        if (bx == n_boxes - 1 && last_cols == 16) tmem_ld16_lo(taddr + bx * 32, r);
        else tmem_ld32(taddr + bx * 32, r);
      };
      // releases: after the loads of positions [0, ovl_boxes) have landed -> shared columns free; after the last -> region free
      auto after_loaded = [&](int i) {
        const bool rel_ovl = (ovl_boxes > 0) ? (i + 1 == ovl_boxes) : (i + 1 == n_boxes);
        const bool rel_all = (i + 1 == n_boxes);
        if (rel_ovl || rel_all) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (rel_ovl) mbar_arrive_leader(smem_u32(tmem_ovl_bar));
            if (rel_all) mbar_arrive_leader(smem_u32(&tmem_empty_bar[par]));
          }
        }
      };
      auto process = [&](const uint32_t (&r)[32], int bx) {
        if (staged) epi_box_staged(p, w, &tmap_o, &tmap_o2, r, bx, row0, valid, lane);
        else epi_box_direct(p, s_add, s_mul, r, bx, bx == n_boxes - 1 ? last_cols : 32, s, in_range, valid, b, sy, sx);
      };
      uint32_t ra[32], rb[32];
      load_box(box_at(0), ra);
      for (int i = 0; i < n_boxes; i += 2) {
        tmem_ld_wait();
        if (i + 1 < n_boxes) load_box(box_at(i + 1), rb);      // in flight while box i is processed
        after_loaded(i);
        process(ra, box_at(i));
        if (i + 1 < n_boxes) {
          tmem_ld_wait();
          if (i + 2 < n_boxes) load_box(box_at(i + 2), ra);
          after_loaded(i + 1);
          process(rb, box_at(i + 1));
        }
      }
      if (w.s_relu) {
        __syncwarp();
        uint32_t* g = p.relu_bits + static_cast<int64_t>(row0) * p.ld_bits;
        const int64_t lim = (p.n_slots - row0) * p.ld_bits;
        for (int i = lane; i < bits_words; i += 32) {
          const int rr = i / p.ld_bits, ww = i - rr * p.ld_bits;
          if (i < lim) g[i] = ww < n_boxes ? w.s_relu[rr * kBitsPitch + ww] : 0u;
        }
        __syncwarp();
      }
    }
    if (staged && lane == 0) bulk_wait_read<0>();            // shared memory must outlive the last TMA store
    if (p.stats && warp == 2 && lane == 0) {
      p.stats[blockIdx.x * 8 + 5] = clock64() - t_begin;     // epilogue total
      p.stats[blockIdx.x * 8 + 6] = t_wait;                  // ... waiting for accumulators
    }
  }

  tc_fence_before();
  __syncthreads();
  if (p.stat_off && p.col_sums) {
    // combine the four epilogue warps' partial sums and add them to the global fp64 accumulators
    const float* st = reinterpret_cast<const float*>(smem + p.stat_off);
    for (int i = threadIdx.x; i < 2 * p.n_pad; i += kConvThreads) {
      const float v = st[i] + st[2 * p.n_pad + i] + st[4 * p.n_pad + i] + st[6 * p.n_pad + i];
      atomicAdd(p.col_sums + i, static_cast<double>(v));
    }
  }
  cluster_sync();                                      // nobody leaves while the peer may still touch its memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// CUDA-core cross-check kernel: one thread per (slot, 16 output channels); fp32 FMA over the same 16-bit operands.
// Supports the plain epilogue (bias / scale+shift / ReLU / gate bits / halo) in all three output modes.
__global__ void conv2x2_simt_kernel(const __nv_bfloat16* __restrict__ in, int ld_in, int cin_pad,
                                    const __nv_bfloat16* __restrict__ wpack, int k_total, const ConvParams p) {
  const int groups = p.n_pad / 16;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t s = idx / groups;
  const int c0 = static_cast<int>(idx - s * groups) * 16;
  if (s >= p.n_slots) return;
  int b, sy, sx;
  slot_coords(p, s, b, sy, sx);
  const bool valid = (p.type == 0 || (sy >= 1 && sx >= 1));
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = 0.f;
  for (int tap = 0; tap < 4; ++tap) {
    const int64_t r = s + p.tap_off[tap];
    if (r < 0 || r >= p.n_slots) continue;
    const __nv_bfloat16* a = in + r * ld_in;
    const __nv_bfloat16* w = wpack + static_cast<int64_t>(tap) * p.n_kc * 64;
    for (int c = 0; c < cin_pad; ++c) {
      const float av = from16(reinterpret_cast<const uint16_t*>(a)[c], p.ab_dtype);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaf(av, from16(reinterpret_cast<const uint16_t*>(w)[static_cast<int64_t>(c0 + j) * k_total + c], p.ab_dtype), v[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int c = c0 + j;
    float mul = 1.f, add = 0.f;
    if (p.has_scale) {
      mul = p.scale[c];
      add = p.shift[c] + (p.bias ? p.bias[c] * mul : 0.f);
    } else if (p.bias) {
      add = p.bias[c];
    }
    float x = epi_value(p, v[j], add, mul, valid);
    if (p.gate_bits && !((p.gate_bits[s * p.ld_bits + (c >> 5)] >> (c & 31)) & 1u)) x = 0.f;
    v[j] = x;
  }
  if (p.out_mode == 0) {
    uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + s * p.ld_out + c0);
    uint4 a, c;
    a.x = pack16x2(v[0], v[1], p.out_dtype);
    a.y = pack16x2(v[2], v[3], p.out_dtype);
    a.z = pack16x2(v[4], v[5], p.out_dtype);
    a.w = pack16x2(v[6], v[7], p.out_dtype);
    c.x = pack16x2(v[8], v[9], p.out_dtype);
    c.y = pack16x2(v[10], v[11], p.out_dtype);
    c.z = pack16x2(v[12], v[13], p.out_dtype);
    c.w = pack16x2(v[14], v[15], p.out_dtype);
    o[0] = a;
    o[1] = c;
  } else {
    direct_store16(p, s, true, valid, b, sy, sx, c0, v);
  }
}

static int fill_params(const mmlf_conv_args* a, ConvParams& p) {
  MMLF_REQUIRE(a != nullptr, "conv2x2: null args");
  MMLF_REQUIRE(a->in && a->wpack && a->out, "conv2x2: null buffer");
  MMLF_REQUIRE(a->cin_pad > 0 && a->cin_pad % 16 == 0, "conv2x2: cin_pad %d must be a positive multiple of 16", a->cin_pad);
  MMLF_REQUIRE(a->n_pad >= 16 && a->n_pad % 16 == 0 && a->n_pad <= kMaxNPad, "conv2x2: n_pad %d must be a multiple of 16 in [16, 320]", a->n_pad);
  MMLF_REQUIRE(a->ld_in % 8 == 0 && a->ld_in >= a->cin_pad, "conv2x2: ld_in %d must be a multiple of 8 and >= cin_pad", a->ld_in);
  MMLF_REQUIRE(a->B > 0 && a->H > 0 && a->W > 0, "conv2x2: bad geometry");
  MMLF_REQUIRE(a->type == 0 || a->type == 1, "conv2x2: type must be 0 or 1");
  MMLF_REQUIRE(a->out_mode >= 0 && a->out_mode <= 2, "conv2x2: bad out_mode");
  MMLF_REQUIRE(a->out_mode == 2 || (a->ld_out >= a->n_pad && a->ld_out % (a->out_mode == 0 ? 8 : 4) == 0),
               "conv2x2: ld_out %d too small / misaligned for n_pad %d", a->ld_out, a->n_pad);
  MMLF_REQUIRE(a->out_mode != 2 || (a->n_real >= 1 && a->n_real <= a->n_pad), "conv2x2: bad n_real");
  MMLF_REQUIRE((a->scale == nullptr) == (a->shift == nullptr), "conv2x2: scale and shift come together");
  MMLF_REQUIRE(!(a->gate_bits || a->relu_bits) || a->ld_bits >= (a->n_pad + 31) / 32,
               "conv2x2: ld_bits %d too small for n_pad %d", a->ld_bits, a->n_pad);
  MMLF_REQUIRE(!(a->gate_bits || a->relu_bits) || a->ld_bits < kBitsPitch, "conv2x2: ld_bits %d too large", a->ld_bits);
  MMLF_REQUIRE(a->out_mode == 0 || !(a->out2 || a->relu_bits || a->col_sums || a->gate_bits),
               "conv2x2: out2 / relu_bits / gate_bits / col_sums need out_mode 0");
  MMLF_REQUIRE(!a->out2 || (a->ld_out2 >= a->n_pad && a->ld_out2 % 8 == 0), "conv2x2: bad ld_out2 %d", a->ld_out2);
  p.Hp = a->H + 1;
  p.Wp = a->W + 1;
  p.H = a->H;
  p.W = a->W;
  p.n_slots = static_cast<int64_t>(a->B) * p.Hp * p.Wp;
  MMLF_REQUIRE(p.n_slots + 2 * kTileM < (1ll << 31), "conv2x2: too many slots for 32-bit TMA coordinates");
  p.n_pad = a->n_pad;
  p.n_parts = a->n_pad > 256 ? 2 : 1;
  p.n_part = a->n_pad / p.n_parts;
  MMLF_REQUIRE(p.n_part % 16 == 0, "conv2x2: n_pad %d does not split into MMA N multiples of 16", a->n_pad);
  p.n_kc = ceil_div(a->cin_pad, 64);
  p.last_ksteps = (a->cin_pad - (p.n_kc - 1) * 64) / 16;
  if (a->type == 0) {
    p.tap_off[0] = 0; p.tap_off[1] = 1; p.tap_off[2] = p.Wp; p.tap_off[3] = p.Wp + 1;
  } else {
    p.tap_off[0] = -p.Wp - 1; p.tap_off[1] = -p.Wp; p.tap_off[2] = -1; p.tap_off[3] = 0;
  }
  p.num_tiles = static_cast<int>(ceil_div64(p.n_slots, 2 * kTileM));
  p.type = a->type;
  p.relu = a->relu;
  p.out_mode = a->out_mode;
  p.n_real = a->n_real;
  p.ld_out = a->ld_out;
  MMLF_REQUIRE((a->ab_dtype | a->out_dtype | a->out2_dtype) >> 1 == 0, "conv2x2: dtype codes are 0 (bf16) or 1 (fp16)");
  p.ab_dtype = a->ab_dtype;
  p.out_dtype = a->out_dtype;
  p.out2_dtype = a->out2_dtype;
  p.has_scale = a->scale != nullptr;
  p.dual = a->out2 != nullptr;
  p.ld_bits = a->ld_bits;
  p.bias = a->bias;
  p.scale = a->scale;
  p.shift = a->shift;
  p.gate_bits = a->gate_bits;
  p.relu_bits = a->relu_bits;
  p.col_sums = a->col_sums;
  p.out = a->out;
  p.stats = nullptr;
  p.stages = 0;
  p.epi_off = p.bits_off = p.stat_off = p.aux_off = 0;
  return 0;
}

}  // namespace mmlf

using namespace mmlf;

static long long* g_conv_stats = nullptr;
// debug hook (not part of the public header): per-CTA cycle counters of the next conv launches, [grid][8] int64
extern "C" void mmlf_debug_conv_stats(long long* device_buf) { g_conv_stats = device_buf; }

extern "C" int mmlf_conv2x2(const mmlf_conv_args* a, void* stream) {
  ConvParams p;
  if (int rc = fill_params(a, p)) return rc;
  p.stats = g_conv_stats;
  const uint32_t stage_bytes = kABytes + p.n_pad * 64;
  // shared-memory plan behind the operand stages: [epilogue staging | bit scratch | statistics | barriers + constants]
  const uint32_t epi_bytes = p.out_mode == 0 ? 4u * (p.dual ? 4u : 2u) * kBoxBytes : 0u;
  const uint32_t bits_bytes = (p.gate_bits || p.relu_bits) ? 4u * 2u * 32u * kBitsPitch * 4u : 0u;
  const uint32_t stat_bytes = p.col_sums ? 4u * 2u * p.n_pad * 4u : 0u;
  const uint32_t aux_bytes = 192 + 2 * kMaxNPad * 4 + 64;   // barriers + TMEM pointer, then the per-channel constants
  const uint32_t tail_bytes = epi_bytes + bits_bytes + ((stat_bytes + 15u) & ~15u) + aux_bytes;
  const uint32_t max_smem = 232448;   // 227 KB opt-in limit per CTA on sm_100
  int stages = static_cast<int>((max_smem - 1024 - tail_bytes) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  MMLF_REQUIRE(stages >= 2, "conv2x2: not enough shared memory for a 2-stage pipeline (n_pad %d)", p.n_pad);
  p.stages = stages;
  p.epi_off = stages * stage_bytes;                       // multiple of 1024 (stage_bytes = 16384 + n_pad * 64)
  p.bits_off = p.epi_off + epi_bytes;
  p.stat_off = stat_bytes ? p.bits_off + bits_bytes : 0;
  p.aux_off = p.bits_off + bits_bytes + ((stat_bytes + 15u) & ~15u);
  const uint32_t smem_bytes = 1024 + p.aux_off + aux_bytes;

  CUtensorMap tmap_a, tmap_b, tmap_o, tmap_o2;
  if (int rc = make_tmap_2d_16(&tmap_a, a->in, a->cin_pad, p.n_slots, static_cast<uint64_t>(a->ld_in) * 2, 64, kTileM, 128))
    return rc;
  const uint64_t k_total = static_cast<uint64_t>(4) * p.n_kc * 64;
  if (int rc = make_tmap_2d_16(&tmap_b, a->wpack, k_total, p.n_pad, k_total * 2, 64, p.n_part / 2, 128)) return rc;
  tmap_o = tmap_a;
  tmap_o2 = tmap_a;
  if (p.out_mode == 0) {
    if (int rc = make_tmap_2d_16(&tmap_o, a->out, p.n_pad, p.n_slots, static_cast<uint64_t>(a->ld_out) * 2, 32, 32, 64))
      return rc;
    if (p.dual)
      if (int rc = make_tmap_2d_16(&tmap_o2, a->out2, p.n_pad, p.n_slots, static_cast<uint64_t>(a->ld_out2) * 2, 32, 32, 64))
        return rc;
  }

  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv2x2_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    MMLF_REQUIRE(e == cudaSuccess, "conv2x2: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int max_pairs = sm_count() / 2;
  const int pairs = p.num_tiles < max_pairs ? p.num_tiles : max_pairs;
  conv2x2_tc2_kernel<<<2 * pairs, kConvThreads, smem_bytes, static_cast<cudaStream_t>(stream)>>>(tmap_a, tmap_b, tmap_o,
                                                                                              tmap_o2, p);
  return check_launch("conv2x2_tc2_kernel");
}

extern "C" int mmlf_conv2x2_simt(const mmlf_conv_args* a, void* stream) {
  ConvParams p;
  if (int rc = fill_params(a, p)) return rc;
  MMLF_REQUIRE(!(a->out2 || a->relu_bits || a->col_sums), "conv2x2_simt: out2 / relu_bits / col_sums are not supported");
  const int64_t total = p.n_slots * (p.n_pad / 16);
  const int threads = 128;
  const int64_t blocks = ceil_div64(total, threads);
  conv2x2_simt_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(a->in), a->ld_in, a->cin_pad,
      reinterpret_cast<const __nv_bfloat16*>(a->wpack), 4 * p.n_kc * 64, p);
  return check_launch("conv2x2_simt_kernel");
}
