// 2x2 convolution as an implicit GEMM on the 5th-gen tensor cores (tcgen05 + TMEM), fed by TMA.
//
// Replaces nn.Conv2d(cin, cout, 2, padding=1|0) of /root/reference/mmlf/model/feed_forward.py:123,125
// (forward) and, with dgrad-packed weights, its data gradient.  See DESIGN.md section 4.
//
//   GEMM view : D[slot][n] = sum_{tap, c} A[slot + off(tap)][c] * Wp[n][tap][c]
//   A operand : activations in the bf16 slot layout; one (tap, 64-channel chunk) = one 2-D TMA box of
//               128 rows x 128 B landing as a canonical K-major SWIZZLE_128B tile.  Rows outside the array and
//               channels >= cin_pad are zero-filled by TMA (that is the conv padding of the first/last image).
//   B operand : packed weights [n_pad][4 * kc * 64], K-major, SWIZZLE_128B, re-streamed from L2 per tile.
//   D         : fp32 in TMEM, 128 lanes x n_pad columns, split into <= 2 MMAs of N <= 256 per k-step.
//   roles     : warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2..5 = epilogue (one TMEM lane
//               quadrant each).  Persistent CTAs, static round-robin over 128-slot tiles.
#include "../../include/mmlf_b200.h"
#include "common.cuh"
#include <stdlib.h>

#include "host_util.h"

namespace mmlf {

constexpr int kTileM = 128;
constexpr int kABytes = kTileM * 128;  // one A stage: 128 rows x 64 bf16
constexpr int kConvThreads = 192;
constexpr int kMaxStages = 8;
constexpr uint32_t kTmemCols = 512;

struct ConvParams {
  int64_t n_slots;
  int Hp, Wp, H, W;
  int n_pad, n_parts, n_part;
  int n_kc, last_ksteps;
  int tap_off[4];
  int num_tiles, stages;
  int b_boxes, b_box_rows;
  int type, relu, out_mode, n_real, ld_out, ld_gate;
  int ab_dtype, gate_dtype, out_dtype;          // 0 = bf16, 1 = fp16
  const float* bias;
  const float* scale;
  const float* shift;
  const __nv_bfloat16* gate;
  void* out;
  long long* stats;                             // optional per-CTA cycle counters (debug): [grid][8]
};

// Epilogue math + store for 16 consecutive output channels [c0, c0+16) of one slot.  Shared by the tensor-core
// kernel and the CUDA-core cross-check kernel so that both have identical semantics.
__device__ __forceinline__ void epilogue_store16(const ConvParams& p, const float* s_bias, const float* s_scale,
                                                 const float* s_shift, int64_t s, bool in_range, bool valid, int b,
                                                 int sy, int sx, int c0, float (&v)[16]) {
  if (!in_range) return;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float x = v[j];
    if (s_bias) x += s_bias[c0 + j];
    if (s_scale) x = x * s_scale[c0 + j] + s_shift[c0 + j];
    if (p.relu) x = fmaxf(x, 0.f);
    v[j] = valid ? x : 0.f;
  }
  if (p.gate) {
    const uint4* g = reinterpret_cast<const uint4*>(p.gate + s * p.ld_gate + c0);
    uint4 g0 = g[0], g1 = g[1];
    const uint32_t gw[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float lo, hi;
      unpack16x2(gw[j], p.gate_dtype, lo, hi);
      if (!(lo > 0.f)) v[2 * j] = 0.f;
      if (!(hi > 0.f)) v[2 * j + 1] = 0.f;
    }
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
  } else if (p.out_mode == 1) {
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

__device__ __forceinline__ void slot_coords(const ConvParams& p, int64_t s, int& b, int& sy, int& sx) {
  const int per_img = p.Hp * p.Wp;
  b = static_cast<int>(s / per_img);
  const int rem = static_cast<int>(s - static_cast<int64_t>(b) * per_img);
  sy = rem / p.Wp;
  sx = rem - sy * p.Wp;
}

__global__ void __launch_bounds__(kConvThreads, 1)
conv2x2_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t tiles_addr = (raw_addr + 1023u) & ~1023u;   // SWIZZLE_128B atoms need 1024 B alignment
  uint8_t* smem = smem_raw + (tiles_addr - raw_addr);
  const uint32_t stage_bytes = kABytes + static_cast<uint32_t>(p.n_pad) * 128u;

  uint8_t* aux = smem + static_cast<size_t>(p.stages) * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full_bar = empty_bar + kMaxStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 1);
  float* s_bias = reinterpret_cast<float*>(tmem_ptr_smem + 4);
  float* s_scale = s_bias + p.n_pad;
  float* s_shift = s_scale + p.n_pad;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < p.stages; ++i) {
        mbar_init(smem_u32(&full_bar[i]), 1);
        mbar_init(smem_u32(&empty_bar[i]), 1);
      }
      mbar_init(smem_u32(tmem_full_bar), 1);
      mbar_init(smem_u32(tmem_empty_bar), 128);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_ptr_smem), kTmemCols);
  }
  for (int i = threadIdx.x; i < p.n_pad; i += kConvThreads) {
    s_bias[i] = p.bias ? p.bias[i] : 0.f;
    s_scale[i] = p.scale ? p.scale[i] : 1.f;
    s_shift[i] = p.shift ? p.shift[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int row0 = tile * kTileM;
        for (int tap = 0; tap < 4; ++tap) {
          for (int kc = 0; kc < p.n_kc; ++kc) {
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
            const uint32_t fb = smem_u32(&full_bar[stage]);
            mbar_arrive_expect_tx(fb, stage_bytes);
            const uint32_t a_dst = tiles_addr + stage * stage_bytes;
            tma_load_2d(a_dst, &tmap_a, fb, kc * 64, row0 + p.tap_off[tap]);
            const uint32_t b_dst = a_dst + kABytes;
            const int kcol = (tap * p.n_kc + kc) * 64;
            for (int bb = 0; bb < p.b_boxes; ++bb)
              tma_load_2d_hint(b_dst + bb * p.b_box_rows * 128, &tmap_b, fb, kcol, bb * p.b_box_rows, kEvictLast);
            if (++stage == static_cast<uint32_t>(p.stages)) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      const uint32_t idesc = make_idesc_16(kTileM, p.n_part, 0, 0, p.ab_dtype, p.ab_dtype);
      uint32_t stage = 0, phase = 0, tphase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(smem_u32(tmem_empty_bar), tphase ^ 1u);   // epilogue has drained the accumulator
        tc_fence_after();
        uint32_t accumulate = 0;
        for (int tap = 0; tap < 4; ++tap) {
          for (int kc = 0; kc < p.n_kc; ++kc) {
            mbar_wait(smem_u32(&full_bar[stage]), phase);
            tc_fence_after();
            const uint32_t a_addr = tiles_addr + stage * stage_bytes;
            const uint32_t b_addr = a_addr + kABytes;
            const int ksteps = (kc == p.n_kc - 1) ? p.last_ksteps : 4;
            for (int k = 0; k < ksteps; ++k) {
              const uint64_t adesc = make_sw128_desc(a_addr + k * 32, 0, 1024);
              for (int part = 0; part < p.n_parts; ++part) {
                const uint64_t bdesc = make_sw128_desc(b_addr + part * p.n_part * 128 + k * 32, 0, 1024);
                umma_f16(tmem_base + part * p.n_part, adesc, bdesc, idesc, accumulate);
              }
              accumulate = 1;
            }
            umma_commit(smem_u32(&empty_bar[stage]));       // frees the smem stage once these MMAs retire
            if (++stage == static_cast<uint32_t>(p.stages)) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
        umma_commit(smem_u32(tmem_full_bar));                // accumulator complete -> epilogue
        tphase ^= 1u;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: TMEM -> regs -> global
    const int q = warp & 3;                                  // TMEM lane quadrant this warp may access
    uint32_t tphase = 0;
    const float* eb = p.bias ? s_bias : nullptr;
    const float* es = p.scale ? s_scale : nullptr;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(smem_u32(tmem_full_bar), tphase);
      tc_fence_after();
      const int64_t s = static_cast<int64_t>(tile) * kTileM + q * 32 + lane;
      const bool in_range = s < p.n_slots;
      int b = 0, sy = 0, sx = 0;
      if (in_range) slot_coords(p, s, b, sy, sx);
      const bool valid = in_range && (p.type == 0 || (sy >= 1 && sx >= 1));
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
      for (int c0 = 0; c0 < p.n_pad; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + c0, r);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        epilogue_store16(p, eb, es, s_shift, s, in_range, valid, b, sy, sx, c0, v);
      }
      tc_fence_before();
      mbar_arrive(smem_u32(tmem_empty_bar));
      tphase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair version (cta_group::2): two SMs of a cluster compute one 256-slot tile.  Each CTA loads its own 128 A rows
// and HALF of the weight rows of every MMA (the hardware reads the other half from the peer's shared memory), which
// halves the dominant shared-memory fill traffic (the weights are re-streamed from L2 for every tile).  The leader CTA
// (cluster rank 0) issues all MMAs; both CTAs run a TMA producer and a 128-row epilogue out of their own TMEM.
//
// TMEM plan: consecutive tiles alternate between two accumulator regions so that the epilogue of tile t overlaps the
// MMAs of tile t+1.  Two 288-column accumulators do not fit into 512 columns, so the regions are [0, n_pad) and
// [512 - n_pad, 512) and share `ovl` = 2 * n_pad - 512 columns; the epilogue drains the shared columns first and
// releases them through a separate barrier, after which the next tile's MMAs may start.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
conv2x2_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                   const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t tiles_addr = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (tiles_addr - raw_addr);
  const uint32_t b_bytes = static_cast<uint32_t>(p.n_pad) * 64u;          // half of the weight rows per CTA
  const uint32_t stage_bytes = kABytes + b_bytes;

  uint8_t* aux = smem + static_cast<size_t>(p.stages) * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full_bar = empty_bar + kMaxStages;      // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;          // [2], leader's copy is used
  uint64_t* tmem_ovl_bar = tmem_empty_bar + 2;           // leader's copy is used
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_ovl_bar + 1);
  float* s_bias = reinterpret_cast<float*>(tmem_ptr_smem + 4);
  float* s_scale = s_bias + p.n_pad;
  float* s_shift = s_scale + p.n_pad;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const bool leader = rank == 0;
  const int base1 = p.n_pad > 256 ? 512 - p.n_pad : 256;      // TMEM column base of odd tiles
  const int ovl = p.n_pad > 256 ? 2 * p.n_pad - 512 : 0;      // columns shared by the two regions (multiple of 32)

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
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
  for (int i = threadIdx.x; i < p.n_pad; i += kConvThreads) {
    s_bias[i] = p.bias ? p.bias[i] : 0.f;
    s_scale[i] = p.scale ? p.scale[i] : 1.f;
    s_shift[i] = p.shift ? p.shift[i] : 0.f;
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
        if (ovl > 0 && it > 0) mbar_wait(smem_u32(tmem_ovl_bar), (it - 1) & 1);   // shared columns drained (tile it - 1)
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
    const int q = warp & 3;
    const float* eb = p.bias ? s_bias : nullptr;
    const float* es = p.scale ? s_scale : nullptr;
    const int n_chunks = p.n_pad >> 4, ovl_chunks = ovl >> 4;
    int it = 0;
    long long t_wait = 0, t_begin = clock64();
    for (int tile = pair; tile < p.num_tiles; tile += n_pairs, ++it) {
      const int par = it & 1;
      const long long tw = clock64();
      mbar_wait(smem_u32(&tmem_full_bar[par]), (it >> 1) & 1);
      t_wait += clock64() - tw;
      tc_fence_after();
      const int64_t s = static_cast<int64_t>(tile) * (2 * kTileM) + rank * kTileM + q * 32 + lane;
      const bool in_range = s < p.n_slots;
      int b = 0, sy = 0, sx = 0;
      if (in_range) slot_coords(p, s, b, sy, sx);
      const bool valid = in_range && (p.type == 0 || (sy >= 1 && sx >= 1));
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (par ? base1 : 0);
      // chunk order: the columns shared with the other region first (the last ones of an even tile, the first ones
      // of an odd tile)
      auto chunk_at = [&](int i) { return (par == 0 && ovl_chunks > 0) ? (i < ovl_chunks ? n_chunks - ovl_chunks + i : i - ovl_chunks) : i; };
      uint32_t ra[16], rb[16];
      float v[16];
      tmem_ld16(taddr + chunk_at(0) * 16, ra);
      for (int i = 0; i < n_chunks; i += 2) {
        tmem_ld_wait();
        if (i + 1 < n_chunks) tmem_ld16(taddr + chunk_at(i + 1) * 16, rb);     // in flight while chunk i is stored
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(ra[j]);
        epilogue_store16(p, eb, es, s_shift, s, in_range, valid, b, sy, sx, chunk_at(i) * 16, v);
        if (i + 1 == ovl_chunks) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(smem_u32(tmem_ovl_bar));
        }
        if (i + 1 < n_chunks) {
          tmem_ld_wait();
          if (i + 2 < n_chunks) tmem_ld16(taddr + chunk_at(i + 2) * 16, ra);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(rb[j]);
          epilogue_store16(p, eb, es, s_shift, s, in_range, valid, b, sy, sx, chunk_at(i + 1) * 16, v);
          if (i + 2 == ovl_chunks) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(smem_u32(tmem_ovl_bar));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (ovl_chunks == 0) mbar_arrive_leader(smem_u32(tmem_ovl_bar));   // keep the phase count in step
        mbar_arrive_leader(smem_u32(&tmem_empty_bar[par]));
      }
    }
    if (p.stats && warp == 2 && lane == 0) {
      p.stats[blockIdx.x * 8 + 5] = clock64() - t_begin;     // epilogue total
      p.stats[blockIdx.x * 8 + 6] = t_wait;                  // ... waiting for accumulators
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();                                      // nobody leaves while the peer may still touch its memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// CUDA-core cross-check kernel: one thread per (slot, 16 output channels); fp32 FMA over the same bf16 operands.
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
  epilogue_store16(p, p.bias, p.scale, p.shift, s, true, valid, b, sy, sx, c0, v);
}

static int fill_params(const mmlf_conv_args* a, ConvParams& p) {
  MMLF_REQUIRE(a != nullptr, "conv2x2: null args");
  MMLF_REQUIRE(a->in && a->wpack && a->out, "conv2x2: null buffer");
  MMLF_REQUIRE(a->cin_pad > 0 && a->cin_pad % 16 == 0, "conv2x2: cin_pad %d must be a positive multiple of 16", a->cin_pad);
  MMLF_REQUIRE(a->n_pad >= 16 && a->n_pad % 16 == 0 && a->n_pad <= 320, "conv2x2: n_pad %d must be a multiple of 16 in [16, 320]", a->n_pad);
  MMLF_REQUIRE(a->ld_in % 8 == 0 && a->ld_in >= a->cin_pad, "conv2x2: ld_in %d must be a multiple of 8 and >= cin_pad", a->ld_in);
  MMLF_REQUIRE(a->B > 0 && a->H > 0 && a->W > 0, "conv2x2: bad geometry");
  MMLF_REQUIRE(a->type == 0 || a->type == 1, "conv2x2: type must be 0 or 1");
  MMLF_REQUIRE(a->out_mode >= 0 && a->out_mode <= 2, "conv2x2: bad out_mode");
  MMLF_REQUIRE(a->out_mode == 2 || (a->ld_out >= a->n_pad && a->ld_out % (a->out_mode == 0 ? 8 : 4) == 0),
               "conv2x2: ld_out %d too small / misaligned for n_pad %d", a->ld_out, a->n_pad);
  MMLF_REQUIRE(a->out_mode != 2 || (a->n_real >= 1 && a->n_real <= a->n_pad), "conv2x2: bad n_real");
  MMLF_REQUIRE((a->scale == nullptr) == (a->shift == nullptr), "conv2x2: scale and shift come together");
  MMLF_REQUIRE(!a->gate || a->ld_gate % 8 == 0, "conv2x2: ld_gate must be a multiple of 8");
  p.Hp = a->H + 1;
  p.Wp = a->W + 1;
  p.H = a->H;
  p.W = a->W;
  p.n_slots = static_cast<int64_t>(a->B) * p.Hp * p.Wp;
  MMLF_REQUIRE(p.n_slots + kTileM < (1ll << 31), "conv2x2: too many slots for 32-bit TMA coordinates");
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
  p.num_tiles = static_cast<int>(ceil_div64(p.n_slots, kTileM));
  p.b_boxes = a->n_pad > 256 ? 2 : 1;
  p.b_box_rows = a->n_pad / p.b_boxes;
  p.type = a->type;
  p.relu = a->relu;
  p.out_mode = a->out_mode;
  p.n_real = a->n_real;
  p.ld_out = a->ld_out;
  p.ld_gate = a->ld_gate;
  MMLF_REQUIRE((a->ab_dtype | a->gate_dtype | a->out_dtype) >> 1 == 0, "conv2x2: dtype codes are 0 (bf16) or 1 (fp16)");
  p.ab_dtype = a->ab_dtype;
  p.gate_dtype = a->gate_dtype;
  p.out_dtype = a->out_dtype;
  p.bias = a->bias;
  p.scale = a->scale;
  p.shift = a->shift;
  p.gate = reinterpret_cast<const __nv_bfloat16*>(a->gate);
  p.out = a->out;
  p.stats = nullptr;
  p.stages = 0;
  return 0;
}

}  // namespace mmlf

using namespace mmlf;

static int conv_impl() {
  // MMLF_CONV_IMPL=1cta selects the single-CTA kernel (debugging); default is the CTA-pair kernel
  static int impl = -1;
  if (impl < 0) {
    const char* e = getenv("MMLF_CONV_IMPL");
    impl = (e && e[0] == '1') ? 1 : 2;
  }
  return impl;
}

static long long* g_conv_stats = nullptr;
// debug hook (not part of the public header): per-CTA cycle counters of the next conv launches, [grid][8] int64
extern "C" void mmlf_debug_conv_stats(long long* device_buf) { g_conv_stats = device_buf; }

extern "C" int mmlf_conv2x2(const mmlf_conv_args* a, void* stream) {
  ConvParams p;
  if (int rc = fill_params(a, p)) return rc;
  p.stats = g_conv_stats;
  const bool pair = conv_impl() == 2;
  const uint32_t stage_bytes = kABytes + (pair ? p.n_pad * 64 : p.n_pad * 128);
  const uint32_t aux_bytes = (2 * kMaxStages + 5) * 8 + 16 + 3 * p.n_pad * 4 + 64;
  const uint32_t max_smem = 232448;   // 227 KB opt-in limit per CTA on sm_100
  int stages = static_cast<int>((max_smem - 1024 - aux_bytes) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  MMLF_REQUIRE(stages >= 2, "conv2x2: not enough shared memory for a 2-stage pipeline (n_pad %d)", p.n_pad);
  p.stages = stages;
  const uint32_t smem_bytes = 1024 + stages * stage_bytes + aux_bytes;

  CUtensorMap tmap_a, tmap_b;
  if (int rc = make_tmap_2d_bf16(&tmap_a, a->in, a->cin_pad, p.n_slots, static_cast<uint64_t>(a->ld_in) * 2, 64, kTileM))
    return rc;
  const uint64_t k_total = static_cast<uint64_t>(4) * p.n_kc * 64;
  const uint32_t b_rows = pair ? p.n_part / 2 : p.b_box_rows;
  if (int rc = make_tmap_2d_bf16(&tmap_b, a->wpack, k_total, p.n_pad, k_total * 2, 64, b_rows)) return rc;

  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv2x2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv2x2_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    MMLF_REQUIRE(e == cudaSuccess, "conv2x2: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  if (pair) {
    p.num_tiles = static_cast<int>(ceil_div64(p.n_slots, 2 * kTileM));
    const int max_pairs = sm_count() / 2;
    const int pairs = p.num_tiles < max_pairs ? p.num_tiles : max_pairs;
    conv2x2_tc2_kernel<<<2 * pairs, kConvThreads, smem_bytes, static_cast<cudaStream_t>(stream)>>>(tmap_a, tmap_b, p);
    return check_launch("conv2x2_tc2_kernel");
  }
  int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  conv2x2_tc_kernel<<<grid, kConvThreads, smem_bytes, static_cast<cudaStream_t>(stream)>>>(tmap_a, tmap_b, p);
  return check_launch("conv2x2_tc_kernel");
}

extern "C" int mmlf_conv2x2_simt(const mmlf_conv_args* a, void* stream) {
  ConvParams p;
  if (int rc = fill_params(a, p)) return rc;
  const int64_t total = p.n_slots * (p.n_pad / 16);
  const int threads = 128;
  const int64_t blocks = ceil_div64(total, threads);
  conv2x2_simt_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(a->in), a->ld_in, a->cin_pad,
      reinterpret_cast<const __nv_bfloat16*>(a->wpack), 4 * p.n_kc * 64, p);
  return check_launch("conv2x2_simt_kernel");
}
