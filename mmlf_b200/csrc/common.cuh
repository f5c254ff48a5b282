// Shared device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM PTX wrappers.
// Hand-written inline PTX (no CUTLASS dependency); encodings follow the PTX ISA for sm_100a.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#ifndef MMLF_WATCHDOG
#define MMLF_WATCHDOG 1     // trap instead of hanging if a pipeline barrier never completes
#endif

namespace mmlf {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Slow path of mbar_wait, kept out of line: every wait site would otherwise carry the whole polling loop with its
// watchdog, and the warp-specialised kernels are sensitive to their instruction-cache footprint.
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
#if MMLF_WATCHDOG
  long long t0 = 0;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xFFFF) == 0) {
      long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) {   // ~2 s at 2 GHz: pipeline is wedged
        printf("mmlf watchdog: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n",
               blockIdx.x, threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
#else
  while (!mbar_try_wait(bar, parity)) {}
#endif
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity);
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2D tiled load global -> shared, completion signalled on an mbarrier (complete_tx::bytes)
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1,
                                                 uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;   // createpolicy-encoded L2 hints (as CUTLASS CacheHintSm90)
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16 inputs with fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---------------------------------------------------------------- CTA pairs (cta_group::2)
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> even CTA of the pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load issued by either CTA of a pair; the transaction bytes are credited to the LEADER CTA's mbarrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1,
                                                 uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// M = 256 MMA across the two SMs of a pair; issued by one thread of the leader CTA
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One pipeline entry worth of pair MMAs in a single asm block: K16 k-steps (K16 = 1..4), each against one or two
// N parts.  The single issuing thread is the bottleneck of the kernel (one 72-cycle MMA must be issued every 72
// cycles), so the descriptors are advanced inside the block with two uniform adds per MMA instead of being rebuilt
// and moved to uniform registers by the compiler for every instruction (measured: ~16 -> ~5 instructions per MMA).
//   d0 / d1  : TMEM addresses of the accumulator parts        a_lo / b0_lo / b1_lo : low words of the descriptors
//   a_hi / desc_hi : high words of the A / the B descriptors       acc : 0 = the very first MMA overwrites the accumulator
//   idesc / idesc1 : instruction descriptors of the two N parts
//   kstep    : descriptor-low-word advance per 16-deep k-step (2 for K-major SWIZZLE_128B, 128 for MN-major)
#define MMLF_MMA_STEP(OFF, PRED, P2)                                            \
  "add.u32 t, %2, " OFF ";\n\t"                                                  \
  "mov.b64 da, {t, %9};\n\t"                                                     \
  "add.u32 t, %3, " OFF ";\n\t"                                                  \
  "mov.b64 db, {t, %5};\n\t"                                                     \
  "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %6, " PRED ";\n\t" P2
#define MMLF_MMA_PART2(OFF, PRED)                                               \
  "add.u32 t, %4, " OFF ";\n\t"                                                  \
  "mov.b64 db, {t, %5};\n\t"                                                     \
  "tcgen05.mma.cta_group::2.kind::f16 [%1], da, db, %8, " PRED ";\n\t"
#define MMLF_MMA_BLOCK(BODY)                                                    \
  asm volatile("{\n\t.reg .pred p, pt;\n\t.reg .b32 t;\n\t.reg .b64 da, db;\n\t" \
               "setp.ne.b32 p, %7, 0;\n\t"                                       \
               "setp.eq.b32 pt, %7, %7;\n\t" BODY "}"                            \
               ::"r"(d0), "r"(d1), "r"(a_lo), "r"(b0_lo), "r"(b1_lo), "r"(desc_hi), "r"(idesc), "r"(acc), "r"(idesc1), "r"(a_hi) \
               : "memory")
template <int kParts, int kStepLo>
__device__ __forceinline__ void umma_f16_pair_entry(uint32_t d0, uint32_t d1, uint32_t a_lo, uint32_t b0_lo,
                                                    uint32_t b1_lo, uint32_t a_hi, uint32_t desc_hi, uint32_t idesc, uint32_t idesc1,
                                                    uint32_t acc, int ksteps) {
  static_assert(kStepLo == 2 || kStepLo == 128, "descriptor advance per k-step");
  if constexpr (kParts == 2 && kStepLo == 2) {
    if (ksteps == 4) {
      MMLF_MMA_BLOCK(MMLF_MMA_STEP("0", "p", MMLF_MMA_PART2("0", "p")) MMLF_MMA_STEP("2", "pt", MMLF_MMA_PART2("2", "pt"))
                     MMLF_MMA_STEP("4", "pt", MMLF_MMA_PART2("4", "pt")) MMLF_MMA_STEP("6", "pt", MMLF_MMA_PART2("6", "pt")));
    } else if (ksteps == 3) {
      MMLF_MMA_BLOCK(MMLF_MMA_STEP("0", "p", MMLF_MMA_PART2("0", "p")) MMLF_MMA_STEP("2", "pt", MMLF_MMA_PART2("2", "pt"))
                     MMLF_MMA_STEP("4", "pt", MMLF_MMA_PART2("4", "pt")));
    } else if (ksteps == 2) {
      MMLF_MMA_BLOCK(MMLF_MMA_STEP("0", "p", MMLF_MMA_PART2("0", "p")) MMLF_MMA_STEP("2", "pt", MMLF_MMA_PART2("2", "pt")));
    } else {
      MMLF_MMA_BLOCK(MMLF_MMA_STEP("0", "p", MMLF_MMA_PART2("0", "p")));
    }
  } else if constexpr (kParts == 1 && kStepLo == 2) {
    if (ksteps == 4) {
      MMLF_MMA_BLOCK(MMLF_MMA_STEP("0", "p", "") MMLF_MMA_STEP("2", "pt", "") MMLF_MMA_STEP("4", "pt", "")
                     MMLF_MMA_STEP("6", "pt", ""));
    } else if (ksteps == 3) {
      MMLF_MMA_BLOCK(MMLF_MMA_STEP("0", "p", "") MMLF_MMA_STEP("2", "pt", "") MMLF_MMA_STEP("4", "pt", ""));
    } else if (ksteps == 2) {
      MMLF_MMA_BLOCK(MMLF_MMA_STEP("0", "p", "") MMLF_MMA_STEP("2", "pt", ""));
    } else {
      MMLF_MMA_BLOCK(MMLF_MMA_STEP("0", "p", ""));
    }
  } else if constexpr (kParts == 2) {      // MN-major operands: 16 slots further along K = 2048 bytes = +128
    MMLF_MMA_BLOCK(MMLF_MMA_STEP("0", "p", MMLF_MMA_PART2("0", "p")) MMLF_MMA_STEP("128", "pt", MMLF_MMA_PART2("128", "pt"))
                   MMLF_MMA_STEP("256", "pt", MMLF_MMA_PART2("256", "pt")) MMLF_MMA_STEP("384", "pt", MMLF_MMA_PART2("384", "pt")));
  } else {
    MMLF_MMA_BLOCK(MMLF_MMA_STEP("0", "p", "") MMLF_MMA_STEP("128", "pt", "") MMLF_MMA_STEP("256", "pt", "")
                   MMLF_MMA_STEP("384", "pt", ""));
  }
}
// arrive on the mbarrier at the same offset in both CTAs of the pair once the issued MMAs have completed
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(static_cast<uint16_t>(3))
      : "memory");
}
// arrive on the LEADER CTA's copy of a barrier from either CTA
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}

// Shared-memory matrix descriptor (sm_100 "version 1"), 128-byte swizzle.
//   bits [0,14)  start address >> 4      bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4 bits [46,48) version = 1     bits [61,64) layout (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
// Instruction descriptor for kind::f16, bf16 x bf16 -> f32.
//   [4,6) D fmt (1 = f32)  [7,10) A fmt (1 = bf16)  [10,13) B fmt  [15] A major  [16] B major (0 = K-major, 1 = MN-major)
//   [17,23) N >> 3         [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_16(uint32_t M, uint32_t N, uint32_t a_mn_major, uint32_t b_mn_major,
                                                    uint32_t a_fp16, uint32_t b_fp16) {
  // A/B format field: 0 = f16, 1 = bf16
  return (1u << 4) | ((a_fp16 ? 0u : 1u) << 7) | ((b_fp16 ? 0u : 1u) << 10) | (a_mn_major << 15) | (b_mn_major << 16) |
         ((N >> 3) << 17) | ((M >> 4) << 24);
}

// TMEM -> registers: 32 lanes x 32 bit, 16 consecutive columns per thread
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- 16-bit storage formats
// dtype codes used across the C ABI: 0 = bf16, 1 = fp16 (IEEE half).  Forward activations and weights default to
// fp16 (11 significant bits, same tensor-core rate as bf16); gradients are bf16 (fp32 exponent range).
constexpr int kBF16 = 0;
constexpr int kFP16 = 1;

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  // saturate instead of producing inf: fp16 tops out at 65504
  lo = fminf(fmaxf(lo, -65504.f), 65504.f);
  hi = fminf(fmaxf(hi, -65504.f), 65504.f);
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack16x2(float lo, float hi, int dtype) {
  return dtype == kFP16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi);
}
__device__ __forceinline__ void unpack16x2(uint32_t v, int dtype, float& lo, float& hi) {
  if (dtype == kFP16) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&v));
    lo = f.x;
    hi = f.y;
  } else {
    lo = bf16_lo(v);
    hi = bf16_hi(v);
  }
}
__device__ __forceinline__ uint16_t to16(float x, int dtype) {
  if (dtype == kFP16) {
    x = fminf(fmaxf(x, -65504.f), 65504.f);
    return __half_as_ushort(__float2half_rn(x));
  }
  return __bfloat16_as_ushort(__float2bfloat16_rn(x));
}
__device__ __forceinline__ float from16(uint16_t v, int dtype) {
  return dtype == kFP16 ? __half2float(__ushort_as_half(v)) : __uint_as_float(static_cast<uint32_t>(v) << 16);
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sum over the thread block (valid in thread 0); red: shared scratch of blockDim.x / 32 doubles
__device__ __forceinline__ double block_sum_double(double v, double* red) {
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) red[wrp] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (blockDim.x >> 5); ++w) r += red[w];
  __syncthreads();
  return r;   // valid in thread 0
}

}  // namespace mmlf
