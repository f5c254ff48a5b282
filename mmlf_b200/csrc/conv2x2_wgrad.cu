// Weight gradient of the 2x2 convolution on tcgen05: both operands are read "MN-major" straight from the
// channel-last slot arrays, so no transposed copies of activations or gradients are ever made.
//
//   dW[(tap, c)][n] = sum_slots act[slot + off(tap)][c] * dout[slot][n]         (autograd of feed_forward.py:123,125)
//
//   GEMM view : M = (tap, 64-channel chunk, channel) -> 4 * kc * 64 rows in blocks of 128 (two chunks),
//               N = n_pad output channels (one TMEM accumulator of <= 320 columns), K = slots.
//   A operand : act rows [64 slots][64 channels] per chunk (TMA box, SWIZZLE_128B), MN-major descriptor
//               (LBO = distance between the two channel chunks, SBO = 1024 B between 8-slot groups).
//   B operand : dout rows [64 slots][64 channels] x ceil(n_pad / 64) boxes, MN-major.
//   split-K   : the slot range is divided over CTAs; partial sums go to a workspace and are reduced by a second,
//               deterministic kernel that also writes the [n][tap][c] layout.
#include "../../include/mmlf_b200.h"
#include "common.cuh"
#include <stdlib.h>

#include "host_util.h"

namespace mmlf {

constexpr int kWgThreads = 192;
constexpr int kWgKb = 64;                       // slots per pipeline stage
constexpr int kWgBox = kWgKb * 128;             // bytes of one [64 slots][64 ch] box
constexpr int kWgMaxStages = 6;

struct WgradParams {
  int64_t n_slots;
  int n_pad, n_boxes;                           // output channels, ceil(n_pad / 64)
  int kc;                                       // channel chunks per tap of the activation operand
  int n_mblocks, ksplits;
  int64_t slots_per_split;                      // multiple of kWgKb
  int tap_off[4];
  int stages;
  int part_n[2], n_parts;
  int act_dtype, dout_dtype;                    // 0 = bf16, 1 = fp16
  float* ws;                                    // [ksplits][n_mblocks * 128][ws_ld]
  int ws_ld;
};

__global__ void __launch_bounds__(kWgThreads, 1)
conv2x2_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_act, const __grid_constant__ CUtensorMap tmap_dout,
                     const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t tiles_addr = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (tiles_addr - raw_addr);
  const uint32_t a_bytes = 2 * kWgBox, b_bytes = static_cast<uint32_t>(p.n_boxes) * kWgBox;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint8_t* aux = smem + static_cast<size_t>(p.stages) * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
  uint64_t* empty_bar = full_bar + kWgMaxStages;
  uint64_t* done_bar = empty_bar + kWgMaxStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mblock = blockIdx.x % p.n_mblocks, ks = blockIdx.x / p.n_mblocks;
  const int64_t k_begin = static_cast<int64_t>(ks) * p.slots_per_split;
  int64_t k_end = k_begin + p.slots_per_split;
  if (k_end > p.n_slots) k_end = p.n_slots;
  const int n_chunks = k_end > k_begin ? static_cast<int>((k_end - k_begin + kWgKb - 1) / kWgKb) : 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_act);
    prefetch_tmap(&tmap_dout);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < p.stages; ++i) {
        mbar_init(smem_u32(&full_bar[i]), 1);
        mbar_init(smem_u32(&empty_bar[i]), 1);
      }
      mbar_init(smem_u32(done_bar), 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_ptr_smem), 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // TMA producer: warp-uniform loop, one elected lane issues (keeps every operand in uniform registers)
    uint32_t stage = 0, phase = 0;
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int row0 = static_cast<int>(k_begin + static_cast<int64_t>(ch) * kWgKb);
      mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
      const uint32_t fb = smem_u32(&full_bar[stage]);
      const uint32_t a_dst = tiles_addr + stage * stage_bytes;
      const uint32_t b_dst = a_dst + a_bytes;
      if (elect_one()) {
        mbar_arrive_expect_tx(fb, stage_bytes);
        for (int h = 0; h < 2; ++h) {
          const int atom = mblock * 2 + h;                 // (tap, chunk) index
          const int tap = atom / p.kc, chunk = atom - tap * p.kc;
          tma_load_2d(a_dst + h * kWgBox, &tmap_act, fb, chunk * 64, row0 + p.tap_off[tap]);
        }
        for (int j = 0; j < p.n_boxes; ++j) tma_load_2d(b_dst + j * kWgBox, &tmap_dout, fb, j * 64, row0);
      }
      __syncwarp();
      if (++stage == static_cast<uint32_t>(p.stages)) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp == 1) {
    // MMA issuer: warp-uniform loop, one elected lane issues MMAs and commits
    const uint32_t idesc0 = make_idesc_16(128, p.part_n[0], 1, 1, p.act_dtype, p.dout_dtype);
    const uint32_t idesc1 = make_idesc_16(128, p.n_parts > 1 ? p.part_n[1] : 16, 1, 1, p.act_dtype, p.dout_dtype);
    const uint32_t part1_off = static_cast<uint32_t>(p.part_n[0] / 64) * kWgBox;
    uint32_t stage = 0, phase = 0, accumulate = 0;
    for (int ch = 0; ch < n_chunks; ++ch) {
      mbar_wait(smem_u32(&full_bar[stage]), phase);
      tc_fence_after();
      const uint32_t a_addr = tiles_addr + stage * stage_bytes;
      const uint32_t b_addr = a_addr + a_bytes;
      const uint64_t adesc0 = make_sw128_desc(a_addr, kWgBox, 1024);
      const uint64_t bdesc0 = make_sw128_desc(b_addr, kWgBox, 1024);
      const uint64_t bdesc1 = make_sw128_desc(b_addr + part1_off, kWgBox, 1024);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kWgKb / 16; ++k) {
          // 16 slots further along K = 2048 bytes = +128 in the (addr >> 4) field
          umma_f16(tmem_base, adesc0 + 128 * k, bdesc0 + 128 * k, idesc0, accumulate);
          if (p.n_parts > 1) umma_f16(tmem_base + p.part_n[0], adesc0 + 128 * k, bdesc1 + 128 * k, idesc1, accumulate);
          accumulate = 1;
        }
        umma_commit(smem_u32(&empty_bar[stage]));
        if (ch == n_chunks - 1) umma_commit(smem_u32(done_bar));
      }
      accumulate = 1;
      __syncwarp();
      if (++stage == static_cast<uint32_t>(p.stages)) {
        stage = 0;
        phase ^= 1u;
      }
    }
    if (n_chunks == 0 && lane == 0) mbar_arrive(smem_u32(done_bar));
  } else {
    const int q = warp & 3;
    mbar_wait(smem_u32(done_bar), 0);
    tc_fence_after();
    const int row = q * 32 + lane;
    float* dst = p.ws + (static_cast<int64_t>(ks) * p.n_mblocks * 128 + mblock * 128 + row) * p.ws_ld;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    for (int c0 = 0; c0 < p.n_pad; c0 += 16) {
      uint32_t r[16];
      if (n_chunks > 0) {
        tmem_ld16(taddr + c0, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = 0u;
      }
      float4* o = reinterpret_cast<float4*>(dst + c0);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        o[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                           __uint_as_float(r[4 * j + 3]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair version (cta_group::2, M = 256): the pair owns one 64-channel chunk of the activation operand for ALL four
// taps.  CTA r loads ONE box of 72 slots x 64 channels starting at the rows of tap (dy = r, dx = 0); tap (r, 1) is the
// same box shifted by one slot row, expressed in the MMA descriptor as a second MN atom 128 B after the first (the
// 128-byte swizzle is a function of the shared-memory address, so overlapping atoms read consistently: measured).
// The gradient operand is split in N between the two CTAs (each supplies half of the columns of every MMA, loaded as
// boxes that START at its slice so both CTAs use identical shared-memory offsets).  Per stage and CTA: 9 KB of
// activations + <= 24 KB of gradients for 64 slots x 256 x 288 MACs, versus 16 + 40 KB for 128 x 288 before.
// two packed 16-bit floats: fp16 -> bf16 (round to nearest even) or bf16 -> fp16
__device__ __forceinline__ uint32_t convert16x2(uint32_t x, bool to_bf16) {
  if (to_bf16) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&x));
    const __nv_bfloat162 b = __float22bfloat162_rn(f);
    return *reinterpret_cast<const uint32_t*>(&b);
  }
  const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&x));
  const __half2 h = __float22half2_rn(f);
  return *reinterpret_cast<const uint32_t*>(&h);
}

constexpr int kWg2ABytes = 72 * 128;
constexpr int kWg2MaxStages = 8;

struct Wgrad2Params {
  int64_t n_slots;
  int n_pad, kc, ksplits;
  int64_t slots_per_split;
  int tap_base[2];                              // row offset of tap (dy, 0), dy = CTA rank
  int stages, group;                            // pipeline stages; 64-slot chunks per stage
  int n_parts, part_n[2], part_col[2];          // MMA N parts: columns [part_col, part_col + part_n)
  int part_box0[2], part_boxes[2];              // per CTA: first B box of the part and number of boxes (1 or 2)
  int nb;                                       // B boxes per CTA and stage
  int act_dtype, dout_dtype;
  float* ws;                                    // [ksplits][4 * kc * 64][ws_ld]
  int ws_ld;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kWgThreads, 1)
conv2x2_wgrad2_kernel(const __grid_constant__ CUtensorMap tmap_act, const __grid_constant__ CUtensorMap tmap_dout,
                      const Wgrad2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t tiles_addr = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (tiles_addr - raw_addr);
  // A stage holds `group` 64-slot chunks: the MMA warp's per-stage cost (barrier wait, fence, election, commit: ~500
  // cycles) is then paid once per 4 * group MMAs instead of once per 4.
  const uint32_t sub_bytes = kWg2ABytes + static_cast<uint32_t>(p.nb) * kWgBox;
  const uint32_t stage_bytes = static_cast<uint32_t>(p.group) * sub_bytes;
  uint8_t* aux = smem + static_cast<size_t>(p.stages) * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
  uint64_t* empty_bar = full_bar + kWg2MaxStages;
  uint64_t* done_bar = empty_bar + kWg2MaxStages;
  uint64_t* afull_bar = done_bar + 1;              // mixed formats: this CTA's activation boxes of a stage have landed
  uint64_t* conv_bar = afull_bar + kWg2MaxStages;  // (leader's copy) both CTAs have converted them: 2 x 4 warps arrive
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(conv_bar + kWg2MaxStages);
  // Mixed operand formats (fp16 activations x bf16 gradients): tcgen05.mma kind::f16 faults on them, and a bf16 copy of
  // every activation written by the forward pass costs a third of the BatchNorm-apply / first-conv store traffic.  So the
  // activation boxes are converted IN PLACE in shared memory by the four warps that otherwise only run the epilogue:
  // TMA -> afull_bar (per CTA) -> convert -> fence.proxy.async -> conv_bar (leader) -> MMA.
  const bool cvt = p.act_dtype != p.dout_dtype;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int chunk = pair % p.kc, ks = pair / p.kc;
  const int64_t k_begin = static_cast<int64_t>(ks) * p.slots_per_split;
  int64_t k_end = k_begin + p.slots_per_split;
  if (k_end > p.n_slots) k_end = p.n_slots;
  const int n_chunks = k_end > k_begin ? static_cast<int>((k_end - k_begin + kWgKb - 1) / kWgKb) : 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_act);
    prefetch_tmap(&tmap_dout);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < p.stages; ++i) {
        mbar_init(smem_u32(&full_bar[i]), 1);
        mbar_init(smem_u32(&empty_bar[i]), 1);
        mbar_init(smem_u32(&afull_bar[i]), 1);
        mbar_init(smem_u32(&conv_bar[i]), 8);
      }
      mbar_init(smem_u32(done_bar), 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_pair(smem_u32(tmem_ptr_smem), 512);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    uint32_t stage = 0, phase = 0;
    for (int ch0 = 0; ch0 < n_chunks; ch0 += p.group) {
      const int n_here = n_chunks - ch0 < p.group ? n_chunks - ch0 : p.group;
      mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
      const uint32_t fb = smem_u32(&full_bar[stage]);
      if (elect_one()) {
        const uint32_t afb = smem_u32(&afull_bar[stage]);
        if (cvt) {
          mbar_arrive_expect_tx(afb, static_cast<uint32_t>(n_here) * kWg2ABytes);
          if (leader) mbar_arrive_expect_tx(fb, 2u * static_cast<uint32_t>(n_here) * (sub_bytes - kWg2ABytes));
        } else if (leader) {
          mbar_arrive_expect_tx(fb, 2u * static_cast<uint32_t>(n_here) * sub_bytes);
        }
        for (int g = 0; g < n_here; ++g) {
          const int row0 = static_cast<int>(k_begin + static_cast<int64_t>(ch0 + g) * kWgKb);
          const uint32_t a_dst = tiles_addr + stage * stage_bytes + g * sub_bytes;
          const uint32_t b_dst = a_dst + kWg2ABytes;
          if (cvt) tma_load_2d_hint(a_dst, &tmap_act, afb, chunk * 64, row0 + p.tap_base[rank], kEvictNormal);
          else tma_load_2d_pair(a_dst, &tmap_act, fb, chunk * 64, row0 + p.tap_base[rank], kEvictNormal);
          for (int part = 0; part < p.n_parts; ++part) {
            const int col = p.part_col[part] + static_cast<int>(rank) * (p.part_n[part] >> 1);
            for (int j = 0; j < p.part_boxes[part]; ++j)
              tma_load_2d_pair(b_dst + (p.part_box0[part] + j) * kWgBox, &tmap_dout, fb, col + 64 * j, row0, kEvictNormal);
          }
        }
      }
      __syncwarp();
      if (++stage == static_cast<uint32_t>(p.stages)) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp == 1) {
    if (leader) {
      const int a_fmt = cvt ? p.dout_dtype : p.act_dtype;          // format of the activation boxes when the MMAs read them
      const uint32_t idesc0 = make_idesc_16(256, p.part_n[0], 1, 1, a_fmt, p.dout_dtype);
      const uint32_t idesc1 = make_idesc_16(256, p.n_parts > 1 ? p.part_n[1] : 16, 1, 1, a_fmt, p.dout_dtype);
      const uint64_t adesc_t = make_sw128_desc(0, 128, 1024), bdesc_t = make_sw128_desc(0, kWgBox, 1024);
      const uint32_t a_lo0 = static_cast<uint32_t>(adesc_t), b_lo0 = static_cast<uint32_t>(bdesc_t);
      const uint32_t desc_hi = static_cast<uint32_t>(adesc_t >> 32);
      uint32_t stage = 0, phase = 0, accumulate = 0;
      for (int ch0 = 0; ch0 < n_chunks; ch0 += p.group) {
        const int n_here = n_chunks - ch0 < p.group ? n_chunks - ch0 : p.group;
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        if (cvt) mbar_wait(smem_u32(&conv_bar[stage]), phase);
        tc_fence_after();
        // warp-uniform chunk loop, election inside (keeps the descriptors in uniform registers)
#pragma unroll 1
        for (int g = 0; g < n_here; ++g) {
          const uint32_t a_addr = tiles_addr + stage * stage_bytes + g * sub_bytes;
          const uint32_t b_addr = a_addr + kWg2ABytes;
          // A: two MN atoms (taps dx = 0, 1) 128 B apart; B: atoms kWgBox apart; 8-slot groups 1024 B apart
          const uint32_t a_lo = a_lo0 + ((a_addr & 0x3FFFFu) >> 4);
          const uint32_t b0_lo = b_lo0 + (((b_addr + p.part_box0[0] * kWgBox) & 0x3FFFFu) >> 4);
          const uint32_t b1_lo = b_lo0 + (((b_addr + p.part_box0[1] * kWgBox) & 0x3FFFFu) >> 4);
          if (elect_one()) {
            if (p.n_parts > 1)
              umma_f16_pair_entry<2, 128>(tmem_base + p.part_col[0], tmem_base + p.part_col[1], a_lo, b0_lo, b1_lo, desc_hi,
                                          desc_hi, idesc0, idesc1, accumulate, 4);
            else
              umma_f16_pair_entry<1, 128>(tmem_base + p.part_col[0], tmem_base, a_lo, b0_lo, b0_lo, desc_hi, desc_hi, idesc0,
                                          idesc0, accumulate, 4);
          }
          accumulate = 1;
          __syncwarp();
        }
        if (elect_one()) {
          umma_commit_pair(smem_u32(&empty_bar[stage]));
          if (ch0 + n_here == n_chunks) umma_commit_pair(smem_u32(done_bar));
        }
        __syncwarp();
        if (++stage == static_cast<uint32_t>(p.stages)) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (n_chunks == 0 && lane == 0) {
        mbar_arrive(smem_u32(done_bar));
        asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(done_bar) | 0x01000000u) : "memory");
      }
    }
  } else {
    if (cvt) {
      const int tid = static_cast<int>(threadIdx.x) - 64;           // 0 .. 127
      const bool to_bf16 = p.dout_dtype == 0;
      uint32_t stage = 0, phase = 0;
      for (int ch0 = 0; ch0 < n_chunks; ch0 += p.group) {
        const int n_here = n_chunks - ch0 < p.group ? n_chunks - ch0 : p.group;
        mbar_wait(smem_u32(&afull_bar[stage]), phase);
        // 576 16-byte pieces per box, 4.5 per thread: all loads of a box in flight before the first conversion (the loop
        // was latency bound with one piece at a time: +12 % on the whole kernel)
#pragma unroll 1
        for (int g = 0; g < n_here; ++g) {
          uint4* tile = reinterpret_cast<uint4*>(smem + static_cast<size_t>(stage) * stage_bytes + static_cast<size_t>(g) * sub_bytes);
          constexpr int kPieces = kWg2ABytes / 16, kPer = (kPieces + 127) / 128;
          uint4 v[kPer];
#pragma unroll
          for (int j = 0; j < kPer; ++j) {
            const int i = tid + 128 * j;
            if (i < kPieces) v[j] = tile[i];
          }
#pragma unroll
          for (int j = 0; j < kPer; ++j) {
            v[j].x = convert16x2(v[j].x, to_bf16);
            v[j].y = convert16x2(v[j].y, to_bf16);
            v[j].z = convert16x2(v[j].z, to_bf16);
            v[j].w = convert16x2(v[j].w, to_bf16);
          }
#pragma unroll
          for (int j = 0; j < kPer; ++j) {
            const int i = tid + 128 * j;
            if (i < kPieces) tile[i] = v[j];
          }
        }
        // generic-proxy writes -> visible to the async proxy (tcgen05.mma reads shared memory through it), then signal
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(smem_u32(&conv_bar[stage]));
        if (++stage == static_cast<uint32_t>(p.stages)) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
    const int q = warp & 3;
    mbar_wait(smem_u32(done_bar), 0);
    tc_fence_after();
    const int row = q * 32 + lane;                        // TMEM lane = M row of this CTA: (dx = row / 64, channel)
    const int tap = 2 * static_cast<int>(rank) + (row >> 6);
    const int m_row = (tap * p.kc + chunk) * 64 + (row & 63);
    float* dst = p.ws + (static_cast<int64_t>(ks) * (4 * p.kc * 64) + m_row) * p.ws_ld;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    for (int c0 = 0; c0 < p.n_pad; c0 += 16) {
      uint32_t r[16];
      if (n_chunks > 0) {
        tmem_ld16(taddr + c0, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = 0u;
      }
      float4* o = reinterpret_cast<float4*>(dst + c0);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        o[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                           __uint_as_float(r[4 * j + 3]));
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// dw[n][tap][c] = sum_ks ws[ks][(tap * kc + c / 64) * 64 + c % 64][n]
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, int ksplits, int m_rows, int ws_ld, int kc, int n_pad,
                                    int cin_pad, float* __restrict__ dw) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<int64_t>(n_pad) * 4 * cin_pad) return;
  // consecutive threads walk n (contiguous in the workspace) for a fixed (tap, c)
  const int n = static_cast<int>(idx % n_pad);
  const int tc = static_cast<int>(idx / n_pad);
  const int tap = tc / cin_pad, c = tc - tap * cin_pad;
  const int m = (tap * kc + (c >> 6)) * 64 + (c & 63);
  float acc = 0.f;
  for (int k = 0; k < ksplits; ++k) acc += ws[(static_cast<int64_t>(k) * m_rows + m) * ws_ld + n];
  dw[(static_cast<int64_t>(n) * 4 + tap) * cin_pad + c] = acc;
}

// Canonical layout of the weight gradient (what autograd hands to the optimizer): dw[cout][cin][2][2], with the
// per-stream tap mapping and the channel-group padding undone (the inverse of mmlf_pack_conv_weight).
struct WgradCanon {
  float* dw;
  int cout, cin, spatial, groups, group_real, group_pad, accumulate;
};

// same reduction as wgrad_reduce_kernel (same order over the K splits), written straight into the canonical tensor.
// One block = 32 output channels x 4 input channels x 4 taps: the partial sums are read along n (contiguous in the
// workspace, a warp per (tap, channel) pair), transposed through shared memory and written as one float4 per (n, channel)
// = the four taps, 64 contiguous bytes per output channel.  (The first version wrote one float per thread at a stride of
// 16 * cin bytes: 19 us per 280 -> 280 layer for 18 MB of L2-resident partial sums, 39 launches per training step.)
constexpr int kRedTN = 32, kRedTC = 4;
__global__ void __launch_bounds__(256) wgrad_reduce_canon_kernel(const float* __restrict__ ws, int ksplits, int m_rows,
                                                                  int ws_ld, int kc, const WgradCanon o) {
  __shared__ __align__(16) float tile[kRedTN][kRedTC * 4 + 4];      // [n][channel * 4 + canonical tap], padded rows
  const int n0 = blockIdx.x * kRedTN, ci0 = blockIdx.y * kRedTC;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int n = n0 + lane;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int pair = wrp + 8 * r;                                   // (tap, local channel)
    const int tap = pair >> 2, cl = pair & 3, ci = ci0 + cl;
    float acc = 0.f;
    if (ci < o.cin && n < o.cout) {
      const int g = ci / o.group_real, c = g * o.group_pad + (ci - g * o.group_real);    // padded channel of the operand
      const int m = (tap * kc + (c >> 6)) * 64 + (c & 63);
      const float* src = ws + static_cast<int64_t>(m) * ws_ld + n;
      const int64_t kstride = static_cast<int64_t>(m_rows) * ws_ld;
      int k = 0;
      for (; k + 4 <= ksplits; k += 4) {                            // four loads in flight, summed in the order of k
        const float v0 = src[(k + 0) * kstride], v1 = src[(k + 1) * kstride], v2 = src[(k + 2) * kstride],
                    v3 = src[(k + 3) * kstride];
        acc += v0; acc += v1; acc += v2; acc += v3;
      }
      for (; k < ksplits; ++k) acc += src[k * kstride];
    }
    // effective tap (p, q) of this stream -> canonical tap (a, b) of w[., ., a, b]  (weights.cu: eff_to_canonical)
    const int p = tap >> 1, q = tap & 1;
    int a, b;
    if (o.spatial == 0) { a = p; b = q; }
    else if (o.spatial == 1) { a = q; b = p; }
    else { a = q; b = 1 - p; }
    tile[lane][cl * 4 + a * 2 + b] = acc;
  }
  __syncthreads();
  if (threadIdx.x < kRedTN * kRedTC) {
    const int nl = threadIdx.x >> 2, cl = threadIdx.x & 3;
    const int nn = n0 + nl, ci = ci0 + cl;
    if (nn < o.cout && ci < o.cin) {
      float4 v = *reinterpret_cast<const float4*>(&tile[nl][cl * 4]);
      float* dst = o.dw + (static_cast<int64_t>(nn) * o.cin + ci) * 4;
      if ((reinterpret_cast<uintptr_t>(o.dw) & 15) == 0) {          // block-uniform: slices of a flat gradient buffer
        float4* d4 = reinterpret_cast<float4*>(dst);                // behind an odd-sized bias are only 4-byte aligned
        if (o.accumulate) {
          const float4 d = *d4;
          v.x = d.x + v.x; v.y = d.y + v.y; v.z = d.z + v.z; v.w = d.w + v.w;
        }
        *d4 = v;
      } else {
        if (o.accumulate) { v.x = dst[0] + v.x; v.y = dst[1] + v.y; v.z = dst[2] + v.z; v.w = dst[3] + v.w; }
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
      }
    }
  }
}

static int launch_wgrad_reduce(const float* ws, int ksplits, int m_rows, int ws_ld, int kc, int n_pad, int cin_pad,
                               float* dw, const WgradCanon* canon, cudaStream_t st) {
  if (canon) {
    const dim3 grid(ceil_div(canon->cout, kRedTN), ceil_div(canon->cin, kRedTC));
    wgrad_reduce_canon_kernel<<<grid, 256, 0, st>>>(ws, ksplits, m_rows, ws_ld, kc, *canon);
    return check_launch("wgrad_reduce_canon_kernel");
  }
  const int64_t total = static_cast<int64_t>(n_pad) * 4 * cin_pad;
  wgrad_reduce_kernel<<<static_cast<unsigned>(ceil_div64(total, 256)), 256, 0, st>>>(ws, ksplits, m_rows, ws_ld, kc, n_pad,
                                                                                     cin_pad, dw);
  return check_launch("wgrad_reduce_kernel");
}

static int wgrad_impl() {
  // MMLF_WGRAD_IMPL=1 selects the single-CTA kernel (debugging); default is the CTA-pair kernel
  static int impl = -1;
  if (impl < 0) {
    const char* e = getenv("MMLF_WGRAD_IMPL");
    impl = (e && e[0] == '1') ? 1 : 2;
  }
  return impl;
}

static void wgrad_shape(int n_pad, int cin_pad, int& kc, int& n_mblocks, int& ksplits, int& ws_ld) {
  kc = ceil_div(cin_pad, 64);
  if (wgrad_impl() == 2) {
    n_mblocks = 2 * kc;                       // rows of the workspace / 128; one CTA pair per 64-channel chunk
    ksplits = (sm_count() / 2) / kc;
  } else {
    n_mblocks = 4 * kc / 2;
    ksplits = sm_count() / n_mblocks;
  }
  if (ksplits < 1) ksplits = 1;
  ws_ld = n_pad;
}

}  // namespace mmlf

using namespace mmlf;

extern "C" int64_t mmlf_conv2x2_wgrad_workspace(int n_pad, int cin_pad) {
  int kc, n_mblocks, ksplits, ws_ld;
  wgrad_shape(n_pad, cin_pad, kc, n_mblocks, ksplits, ws_ld);
  return static_cast<int64_t>(ksplits) * n_mblocks * 128 * ws_ld * sizeof(float);
}

static int wgrad_pair(const void* dout, int ld_dout, int n_pad, const void* act, int ld_act, int cin_pad, int B, int H,
                      int W, int type, int act_dtype, int dout_dtype, float* workspace, float* dw, const WgradCanon* canon,
                      void* stream) {
  Wgrad2Params p;
  const int Wp = W + 1;
  p.n_slots = static_cast<int64_t>(B) * (H + 1) * Wp;
  MMLF_REQUIRE(p.n_slots + 4096 < (1ll << 31), "wgrad: too many slots");
  p.n_pad = n_pad;
  int n_mblocks;
  wgrad_shape(n_pad, cin_pad, p.kc, n_mblocks, p.ksplits, p.ws_ld);
  int64_t per = ceil_div64(p.n_slots, p.ksplits);
  per = ceil_div64(per, kWgKb) * kWgKb;
  p.slots_per_split = per;
  if (type == 0) {
    p.tap_base[0] = 0; p.tap_base[1] = Wp;
  } else {
    p.tap_base[0] = -Wp - 1; p.tap_base[1] = -1;
  }
  // N parts: <= 256 columns each; every CTA supplies half of a part as 1 or 2 boxes of 64 columns starting at its slice
  p.n_parts = n_pad > 256 ? 2 : 1;
  p.part_col[0] = 0;
  p.part_n[0] = n_pad > 256 ? 256 : n_pad;
  p.part_col[1] = 256;
  p.part_n[1] = n_pad > 256 ? n_pad - 256 : 0;
  p.nb = 0;
  for (int i = 0; i < 2; ++i) {
    p.part_box0[i] = p.nb;
    p.part_boxes[i] = i < p.n_parts ? ceil_div(p.part_n[i] / 2, 64) : 0;
    p.nb += p.part_boxes[i];
    MMLF_REQUIRE(i >= p.n_parts || p.part_n[i] % 16 == 0, "wgrad: N part %d is not a multiple of 16", p.part_n[i]);
  }
  p.ws = workspace;
  p.act_dtype = act_dtype;
  p.dout_dtype = dout_dtype;
  const uint32_t sub_bytes = kWg2ABytes + p.nb * kWgBox;
  const uint32_t aux_bytes = (4 * kWg2MaxStages + 1) * 8 + 16 + 64;
  const uint32_t max_smem = 232448;
  // chunks per stage: as many as leave room for a 3-stage pipeline, at most 4
  int group = static_cast<int>((max_smem - 1024 - aux_bytes) / (3 * sub_bytes));
  if (group > 4) group = 4;
  if (group < 1) group = 1;
  {
    static int forced = -1;
    if (forced < 0) {
      const char* e = getenv("MMLF_WGRAD_GROUP");
      forced = e ? atoi(e) : 0;
    }
    if (forced > 0) group = forced;
  }
  p.group = group;
  const uint32_t stage_bytes = group * sub_bytes;
  int stages = static_cast<int>((max_smem - 1024 - aux_bytes) / stage_bytes);
  if (stages > kWg2MaxStages) stages = kWg2MaxStages;
  MMLF_REQUIRE(stages >= 2, "wgrad: not enough shared memory");
  p.stages = stages;
  const uint32_t smem_bytes = 1024 + stages * stage_bytes + aux_bytes;

  CUtensorMap tmap_act, tmap_dout;
  if (int rc = make_tmap_2d_16(&tmap_act, act, cin_pad, p.n_slots, static_cast<uint64_t>(ld_act) * 2, 64, 72, 128)) return rc;
  if (int rc = make_tmap_2d_16(&tmap_dout, dout, n_pad, p.n_slots, static_cast<uint64_t>(ld_dout) * 2, 64, kWgKb, 128)) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv2x2_wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    MMLF_REQUIRE(e == cudaSuccess, "wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  conv2x2_wgrad2_kernel<<<2 * p.kc * p.ksplits, kWgThreads, smem_bytes, st>>>(tmap_act, tmap_dout, p);
  if (int rc = check_launch("conv2x2_wgrad2_kernel")) return rc;
  return launch_wgrad_reduce(workspace, p.ksplits, 4 * p.kc * 64, p.ws_ld, p.kc, n_pad, cin_pad, dw, canon, st);
}

static int wgrad_common(const void* dout, int ld_dout, int n_pad, const void* act, int ld_act, int cin_pad, int B, int H,
                        int W, int type, int act_dtype, int dout_dtype, float* workspace, float* dw,
                        const WgradCanon* canon, void* stream) {
  MMLF_REQUIRE(dout && act && workspace && (dw || canon), "wgrad: null buffer");
  MMLF_REQUIRE((act_dtype | dout_dtype) >> 1 == 0, "wgrad: dtype codes are 0 (bf16) or 1 (fp16)");
  MMLF_REQUIRE(act_dtype == dout_dtype || wgrad_impl() == 2,
               "wgrad: the single-CTA debugging kernel needs both operands in one 16-bit format");
  MMLF_REQUIRE(n_pad % 16 == 0 && n_pad >= 16 && n_pad <= 320, "wgrad: n_pad %d must be a multiple of 16 in [16, 320]", n_pad);
  MMLF_REQUIRE(cin_pad % 16 == 0 && cin_pad >= 16 && cin_pad <= 320, "wgrad: cin_pad %d must be a multiple of 16 in [16, 320]", cin_pad);
  MMLF_REQUIRE(ld_dout % 8 == 0 && ld_act % 8 == 0 && ld_dout >= n_pad && ld_act >= cin_pad, "wgrad: bad row pitch");
  MMLF_REQUIRE(type == 0 || type == 1, "wgrad: type must be 0 or 1");
  if (wgrad_impl() == 2)
    return wgrad_pair(dout, ld_dout, n_pad, act, ld_act, cin_pad, B, H, W, type, act_dtype, dout_dtype, workspace, dw, canon,
                      stream);
  WgradParams p;
  const int Hp = H + 1, Wp = W + 1;
  p.n_slots = static_cast<int64_t>(B) * Hp * Wp;
  MMLF_REQUIRE(p.n_slots + 4096 < (1ll << 31), "wgrad: too many slots");
  p.n_pad = n_pad;
  p.n_boxes = ceil_div(n_pad, 64);
  wgrad_shape(n_pad, cin_pad, p.kc, p.n_mblocks, p.ksplits, p.ws_ld);
  int64_t per = ceil_div64(p.n_slots, p.ksplits);
  per = ceil_div64(per, kWgKb) * kWgKb;
  p.slots_per_split = per;
  if (type == 0) {
    p.tap_off[0] = 0; p.tap_off[1] = 1; p.tap_off[2] = Wp; p.tap_off[3] = Wp + 1;
  } else {
    p.tap_off[0] = -Wp - 1; p.tap_off[1] = -Wp; p.tap_off[2] = -1; p.tap_off[3] = 0;
  }
  if (n_pad > 256) {
    p.n_parts = 2; p.part_n[0] = 256; p.part_n[1] = n_pad - 256;
  } else {
    p.n_parts = 1; p.part_n[0] = n_pad; p.part_n[1] = 0;
  }
  p.ws = workspace;
  p.act_dtype = act_dtype;
  p.dout_dtype = dout_dtype;
  const uint32_t stage_bytes = (2 + p.n_boxes) * kWgBox;
  const uint32_t aux_bytes = (2 * kWgMaxStages + 1) * 8 + 16 + 64;
  const uint32_t max_smem = 232448;
  int stages = static_cast<int>((max_smem - 1024 - aux_bytes) / stage_bytes);
  if (stages > kWgMaxStages) stages = kWgMaxStages;
  MMLF_REQUIRE(stages >= 2, "wgrad: not enough shared memory");
  p.stages = stages;
  const uint32_t smem_bytes = 1024 + stages * stage_bytes + aux_bytes;

  CUtensorMap tmap_act, tmap_dout;
  if (int rc = make_tmap_2d_16(&tmap_act, act, cin_pad, p.n_slots, static_cast<uint64_t>(ld_act) * 2, 64, kWgKb, 128)) return rc;
  if (int rc = make_tmap_2d_16(&tmap_dout, dout, n_pad, p.n_slots, static_cast<uint64_t>(ld_dout) * 2, 64, kWgKb, 128)) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv2x2_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    MMLF_REQUIRE(e == cudaSuccess, "wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  conv2x2_wgrad_kernel<<<p.n_mblocks * p.ksplits, kWgThreads, smem_bytes, st>>>(tmap_act, tmap_dout, p);
  if (int rc = check_launch("conv2x2_wgrad_kernel")) return rc;
  return launch_wgrad_reduce(workspace, p.ksplits, p.n_mblocks * 128, p.ws_ld, p.kc, n_pad, cin_pad, dw, canon, st);
}

extern "C" int mmlf_conv2x2_wgrad(const void* dout, int ld_dout, int n_pad, const void* act, int ld_act, int cin_pad,
                                  int B, int H, int W, int type, int act_dtype, int dout_dtype, float* workspace, float* dw,
                                  void* stream) {
  MMLF_REQUIRE(dw, "wgrad: null buffer");
  return wgrad_common(dout, ld_dout, n_pad, act, ld_act, cin_pad, B, H, W, type, act_dtype, dout_dtype, workspace, dw,
                      nullptr, stream);
}

extern "C" int mmlf_conv2x2_wgrad_canonical(const void* dout, int ld_dout, int n_pad, const void* act, int ld_act,
                                            int cin_pad, int B, int H, int W, int type, int act_dtype, int dout_dtype,
                                            float* workspace, int cout, int cin, int spatial, int in_groups, int group_real,
                                            int group_pad, float* dw, int accumulate, void* stream) {
  MMLF_REQUIRE(dw, "wgrad_canonical: null buffer");
  MMLF_REQUIRE(spatial >= 0 && spatial <= 2, "wgrad_canonical: spatial must be 0..2");
  MMLF_REQUIRE(in_groups >= 1 && group_real >= 1 && group_pad >= group_real && in_groups * group_real == cin &&
                   n_pad >= cout && cin_pad >= in_groups * group_pad,
               "wgrad_canonical: inconsistent channel layout");
  const WgradCanon canon{dw, cout, cin, spatial, in_groups, group_real, group_pad, accumulate};
  return wgrad_common(dout, ld_dout, n_pad, act, ld_act, cin_pad, B, H, W, type, act_dtype, dout_dtype, workspace, nullptr,
                      &canon, stream);
}
