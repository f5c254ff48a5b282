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
//               per-warp swizzled shared-memory staging (32 slots x 64 channels) -> coalesced 128-bit global stores
//               (TMA stores queue behind the producer's bulk loads and cost > 1000 cycles each: measured); optional
//               second copy in the other 16-bit format, ReLU sign bits, per-channel sum / sum of squares.
#include "../../include/mmlf_b200.h"
#include "common.cuh"
#include <stdlib.h>

#include "host_util.h"

namespace mmlf {

constexpr int kTileM = 128;
constexpr int kABoxRows = kTileM + 8;   // one activation box serves both dx taps of a dy: rows [r, r + 128] (+7 to keep
                                        // the 8-row swizzle atoms whole)
constexpr int kABytes = kABoxRows * 128;  // one A stage: 136 rows x 64 16-bit channels
constexpr int kEpiWarps = 8;           // two per TMEM lane quadrant
constexpr int kConvThreads = 64 + 32 * kEpiWarps;
constexpr int kMaxStages = 8;
constexpr uint32_t kTmemCols = 512;
constexpr int kSegBytes = 32 * 64;     // epilogue staging per warp: 32 slots x 32 channels x 2 B, 16-byte chunks XOR-swizzled
constexpr int kMaxNPad = 320;

struct ConvParams {
  int64_t n_slots;
  int Hp, Wp, H, W;
  int n_pad, n_parts, n_part;
  int n_kc, last_ksteps;                        // K chunks per tap (all terms) / 16-deep steps of a term's last chunk
  int kc_term, lo_col;                          // chunks per split term (= n_kc unless split_in), column of the lo block
  int tap_off[4];
  int num_tiles, stages;
  int pair_issue;                               // narrow layers: the MMA warp issues two pipeline stages per round
  int type, relu, out_mode, n_real, ld_out, ld_out2;
  int ab_dtype, out_dtype, out2_dtype;
  int has_scale, dual, ld_bits, split_out;
  uint32_t epi_off, stat_off, aux_off;             // shared-memory offsets from the 1024-aligned base
  const float* bias;
  const float* scale;
  const float* shift;
  const uint32_t* gate_bits;
  uint32_t* relu_bits;
  double* col_sums;
  const uint16_t* bn_z;                         // BatchNorm-backward statistics (kFBnBwd): pre-activation z of the BN layer
  int ld_z, z_dtype;
  const float* bn_scale;
  const float* bn_shift;
  const float* bn_mean;
  void* out;
  void* out2;
  long long* stats;                             // optional per-CTA cycle counters (debug): [grid][16]
};

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

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
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// Epilogue feature mask.  The kernel is instantiated for the combinations the network uses (every runtime branch in
// the epilogue costs issue slots that the lone epilogue warps of an SM sub-partition cannot hide) plus one generic
// instance (kFGeneric) that reads all switches from the parameter block.
enum : uint32_t {
  kFScale = 1,      // y = acc * scale + shift instead of acc + bias
  kFRelu = 2,
  kFGate = 4,       // multiply by the saved ReLU sign bits
  kFBits = 8,       // emit ReLU sign bits
  kFDual = 16,      // second 16-bit copy of the output
  kFStats = 32,     // per-channel sum / sum of squares
  kFOutF16 = 64,    // primary output is fp16 (else bf16)
  kFOut2F16 = 128,  // secondary output is fp16 (else bf16)
  kFDirect = 256,   // out_mode 1 / 2: fp32 rows or planar fp32, written straight from registers
  kFGeneric = 512,
  kFSplit = 1024,   // second output = fp16 residual of the fp16 rounding of the first (split-precision inference)
  kFBnBwd = 2048,   // column statistics are BatchNorm-backward sums (sum g*m, sum g*m*(z - mean)) instead of (sum, sum sq)
};

template <uint32_t F> struct Flags {
  static constexpr bool generic = (F & kFGeneric) != 0;
  __device__ __forceinline__ static bool scale(const ConvParams& p) { return generic ? p.has_scale != 0 : (F & kFScale) != 0; }
  __device__ __forceinline__ static bool relu(const ConvParams& p) { return generic ? p.relu != 0 : (F & kFRelu) != 0; }
  __device__ __forceinline__ static bool gate(const ConvParams& p) { return generic ? p.gate_bits != nullptr : (F & kFGate) != 0; }
  __device__ __forceinline__ static bool bits(const ConvParams& p) { return generic ? p.relu_bits != nullptr : (F & kFBits) != 0; }
  __device__ __forceinline__ static bool dual(const ConvParams& p) { return generic ? p.dual != 0 : (F & kFDual) != 0; }
  __device__ __forceinline__ static bool stats(const ConvParams& p) { return generic ? p.col_sums != nullptr : (F & kFStats) != 0; }
  __device__ __forceinline__ static bool out_f16(const ConvParams& p) { return generic ? p.out_dtype == kFP16 : (F & kFOutF16) != 0; }
  __device__ __forceinline__ static bool out2_f16(const ConvParams& p) { return generic ? p.out2_dtype == kFP16 : (F & kFOut2F16) != 0; }
  __device__ __forceinline__ static bool direct(const ConvParams& p) { return generic ? p.out_mode != 0 : (F & kFDirect) != 0; }
  __device__ __forceinline__ static bool split(const ConvParams& p) { return generic ? p.split_out != 0 : (F & kFSplit) != 0; }
  __device__ __forceinline__ static bool bnbwd(const ConvParams& p) { return generic ? p.bn_z != nullptr : (F & kFBnBwd) != 0; }
};

// One epilogue warp's context for the staged (out_mode 0) path
struct EpiWarp {
  uint32_t stg_addr;        // shared address of this warp's staging rows: [1 + dual] x kSegBytes
  uint32_t zstg_addr;       // staging rows of the BatchNorm pre-activation z (kFBnBwd), same layout
  float* s_stat;            // [2][n_pad] partial column sums of this warp's TMEM lane quadrant (or nullptr)
  const float* s_add;
  const float* s_mul;
};

// One segment of up to 32 output channels [c0, c0 + ncols) of the warp's 32 slots (c0 is a multiple of 32): epilogue
// math, ReLU bits, 16-bit packing into the staging rows, coalesced write-out, column statistics.
// r holds the fp32 accumulators of this thread's slot.  gate_word: this slot's saved ReLU bits of channels [c0, c0+32).
template <uint32_t F>
__device__ __forceinline__ void epi_segment(const ConvParams& p, const EpiWarp& w, const uint32_t (&r)[32], int c0,
                                            int ncols, uint32_t gate_word, bool valid, int row0, int rows_valid,
                                            int lane) {
  using FL = Flags<F>;
  // BatchNorm-backward statistics: this slot's z values of the segment (64 bytes), requested before anything else so that
  // the loads overlap the epilogue math
  uint4 zrow[4];
  if (FL::bnbwd(p)) {
    const int nch = ncols >> 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      zrow[j] = make_uint4(0u, 0u, 0u, 0u);
      if (j < nch && lane < rows_valid)
        zrow[j] = __ldg(reinterpret_cast<const uint4*>(p.bn_z + (static_cast<int64_t>(row0) + lane) * p.ld_z + c0) + j);
    }
  }
  float x[32];
  {
    const float4* add4 = reinterpret_cast<const float4*>(w.s_add + c0);
    if (FL::scale(p)) {
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
  if (FL::relu(p)) {
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
  }
  if (FL::gate(p)) {
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = ((gate_word >> j) & 1u) ? x[j] : 0.f;
  }
  const uint32_t vmask = valid ? 0xFFFFFFFFu : 0u;      // halo slots and slots past the end store zeros
  uint32_t bits = 0;
  if (FL::bits(p)) {
#pragma unroll
    for (int j = 0; j < 32; ++j) bits |= (x[j] > 0.f ? 1u : 0u) << j;
    if (ncols < 32) bits &= 0xFFFFu;
    bits &= vmask;
    if (lane < rows_valid) p.relu_bits[(static_cast<int64_t>(row0) + lane) * p.ld_bits + (c0 >> 5)] = bits;
  }
  // slot row `lane` occupies 64 B of the staging buffer; 16-byte chunk j lands at chunk j ^ ((lane >> 1) & 3)
  const uint32_t row_addr = w.stg_addr + lane * 64;
  const uint32_t sw = (lane >> 1) & 3;
  const int nch = ncols >> 3;
  __syncwarp();                                           // the previous segment's write-out has finished reading
#pragma unroll
  for (int o = 0; o < 2; ++o) {
    if (o == 1 && !FL::dual(p)) break;
    const bool f16 = o ? FL::out2_f16(p) : FL::out_f16(p);
    uint32_t pk[16];
    if (o == 1 && FL::split(p)) {
      // lo = fp16(x - float(fp16(x))): the part of x the first output lost
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint32_t hi2 = pack_f16x2_sat(x[2 * j], x[2 * j + 1]);
        const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi2));
        pk[j] = pack_f16x2_sat(x[2 * j] - hf.x, x[2 * j + 1] - hf.y) & vmask;
      }
    } else if (f16) {
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = pack_f16x2_sat(x[2 * j], x[2 * j + 1]) & vmask;
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(x[2 * j], x[2 * j + 1]) & vmask;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < nch) sts128(row_addr + o * kSegBytes + ((j ^ sw) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
  }
  if (FL::bnbwd(p)) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < nch) sts128(w.zstg_addr + lane * 64 + ((j ^ sw) << 4), zrow[j].x, zrow[j].y, zrow[j].z, zrow[j].w);
  }
  __syncwarp();
  // write-out: 4 lanes cover the 64 staged bytes of one slot row, one store instruction covers 8 rows
#pragma unroll
  for (int o = 0; o < 2; ++o) {
    if (o == 1 && !FL::dual(p)) break;
    uint16_t* gbase = reinterpret_cast<uint16_t*>(o ? p.out2 : p.out);
    const int ld = o ? p.ld_out2 : p.ld_out;
    const uint32_t sbase = w.stg_addr + o * kSegBytes;
    if (ncols == 32) {
      const int r0 = lane >> 2, ch = lane & 3;
      const uint32_t a0 = sbase + r0 * 64 + ((ch ^ ((r0 >> 1) & 3)) << 4);
      uint16_t* g = gbase + (static_cast<int64_t>(row0) + r0) * ld + c0 + ch * 8;
      const int64_t gstep = static_cast<int64_t>(8) * ld;
      if (rows_valid == 32) {
#pragma unroll
        for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(g + k * gstep) = lds128(a0 + k * 512);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (8 * k + r0 < rows_valid) *reinterpret_cast<uint4*>(g + k * gstep) = lds128(a0 + k * 512);
      }
    } else {                                              // 16 channels: 2 lanes per row, 16 rows per instruction
      const int r0 = lane >> 1, ch = lane & 1;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int row = 16 * k + r0;
        const uint4 v = lds128(sbase + row * 64 + ((ch ^ ((row >> 1) & 3)) << 4));
        if (row < rows_valid) *reinterpret_cast<uint4*>(gbase + (static_cast<int64_t>(row0) + row) * ld + c0 + ch * 8) = v;
      }
    }
  }
  if (FL::stats(p)) {
    // lane owns channel c0 + lane over the 32 slots (2-byte reads of one row are conflict-free)
    if (lane < ncols) {
      float s0 = 0.f, q0 = 0.f;
      const uint32_t col = w.stg_addr + (lane & 7) * 2;
      const uint32_t chunk = lane >> 3;
      if (FL::bnbwd(p)) {
        // BatchNorm-backward sums: the staged values are g = d loss / d y of y = relu(z * sc + sh); z of the same slots
        // was staged next to them: accumulate g*m and g*m*(z - mean), m = (z * sc + sh > 0)
        const int c = c0 + lane;
        const float sc = __ldg(p.bn_scale + c), sh = __ldg(p.bn_shift + c), mu = __ldg(p.bn_mean + c);
        const uint32_t zcol = w.zstg_addr + (lane & 7) * 2;
#pragma unroll 8
        for (int row = 0; row < 32; ++row) {
          const uint32_t off = row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4);
          uint16_t v, zr;
          asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(col + off) : "memory");
          asm volatile("ld.shared.u16 %0, [%1];" : "=h"(zr) : "r"(zcol + off) : "memory");
          const float f = FL::out_f16(p) ? __half2float(__ushort_as_half(v)) : __uint_as_float(static_cast<uint32_t>(v) << 16);
          const float zf = p.z_dtype == kFP16 ? __half2float(__ushort_as_half(zr)) : __uint_as_float(static_cast<uint32_t>(zr) << 16);
          const float g = fmaf(zf, sc, sh) > 0.f ? f : 0.f;
          s0 += g;
          q0 = fmaf(g, zf - mu, q0);
        }
      } else {
#pragma unroll 8
        for (int row = 0; row < 32; ++row) {
          uint16_t v;
          asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(col + row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4)) : "memory");
          const float f = FL::out_f16(p) ? __half2float(__ushort_as_half(v)) : __uint_as_float(static_cast<uint32_t>(v) << 16);
          s0 += f;
          q0 = fmaf(f, f, q0);
        }
      }
      w.s_stat[c0 + lane] += s0;
      w.s_stat[p.n_pad + c0 + lane] += q0;
    }
  }
}

// Direct-store variant (fp32 slot rows / planar fp32 outputs of the heads)
__device__ __forceinline__ void epi_box_direct(const ConvParams& p, const float* s_add, const float* s_mul,
                                               const uint32_t (&r)[32], int c0, int ncols, int64_t s, bool in_range,
                                               bool valid, int b, int sy, int sx) {
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
//
// Epilogue warps: two per TMEM lane quadrant (warps 2..5 take the even 32-column segments of a tile, warps 6..9 the odd
// ones), i.e. two per SM sub-partition, so that one warp's shared-memory / TMEM latencies are covered by the other.
template <uint32_t F>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
conv2x2_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                   const ConvParams p) {
  using FL = Flags<F>;
  extern __shared__ uint8_t smem_raw[];
  if (p.stats && threadIdx.x == 0) p.stats[blockIdx.x * 16 + 8] = static_cast<long long>(global_ns());   // kernel entry
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t tiles_addr = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (tiles_addr - raw_addr);
  const uint32_t b_bytes = static_cast<uint32_t>(p.n_pad) * 64u;          // half of the weight rows of one tap per CTA
  const uint32_t stage_bytes = kABytes + 2u * b_bytes;                    // one (dy, 64-channel chunk): A box + both dx taps

  uint8_t* aux = smem + p.aux_off;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full_bar = empty_bar + kMaxStages;      // [4]
  uint64_t* tmem_empty_bar = tmem_full_bar + 4;          // [4], leader's copy is used
  uint64_t* tmem_ovl_bar = tmem_empty_bar + 4;           // leader's copy is used
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_ovl_bar + 1);
  float* s_add = reinterpret_cast<float*>(aux + 256);   // 16-byte aligned (read as float4)
  float* s_mul = s_add + kMaxNPad;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const bool leader = rank == 0;
  // the overlap trick needs the shared columns to be whole 32-channel segments; otherwise the regions are used
  // strictly one after the other
  const int ovl_cols = p.n_pad > 256 ? 2 * p.n_pad - 512 : 0;
  const bool ovl_ok = (ovl_cols & 31) == 0 && (p.n_pad & 31) == 0;
  const int base1 = p.n_pad > 256 ? 512 - p.n_pad : 256;      // TMEM column base of odd tiles
  const int ovl = ovl_ok ? ovl_cols : 0;
  // Narrow layers (n_pad <= 128) are bound by the epilogue, not by the MMAs: four 128-column accumulator regions, and
  // the two epilogue warps of a TMEM lane quadrant take alternate TILES (all segments) instead of alternate segments of
  // the same tile, so two tiles drain concurrently and the 32/32/16-column split no longer leaves one warp with 2/3 of
  // the work.
  const bool quad = p.n_pad <= 128;

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
      for (int i = 0; i < 4; ++i) {
        mbar_init(smem_u32(&tmem_full_bar[i]), 1);
        // one lane of every epilogue warp of both CTAs that works on a tile (half of them in the 4-region scheme)
        mbar_init(smem_u32(&tmem_empty_bar[i]), quad ? kEpiWarps : 2 * kEpiWarps);
      }
      mbar_init(smem_u32(tmem_ovl_bar), 2 * kEpiWarps);
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
    for (int i = threadIdx.x; i < kEpiWarps * 2 * p.n_pad; i += kConvThreads) st[i] = 0.f;
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
    const bool prof = p.stats != nullptr;
    uint32_t stage = 0, phase = 0;
    long long t_wait = 0, t_begin = prof ? clock64() : 0;
    unsigned long long ns_begin = 0;
    if (prof) {
      ns_begin = global_ns();
      if (lane == 0) p.stats[blockIdx.x * 16 + 9] = static_cast<long long>(ns_begin);          // main loop entry
    }
    const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
    const int b_row0 = static_cast<int>(rank) * half_rows;
    const uint32_t tx_bytes = 2 * stage_bytes;
    if (p.pair_issue) {
      // ---- narrow layers: two stages per round, mirroring the MMA warp (both chunks of a dy, or both dy of a tile)
      const int off_dy0 = p.tap_off[0], off_dy1 = p.tap_off[2];
      const int n_kc = p.n_kc, rounds = p.n_kc;
      const uint32_t n_stages = static_cast<uint32_t>(p.stages);
      for (int tile = pair; tile < p.num_tiles; tile += n_pairs) {
        const int row0 = tile * (2 * kTileM) + static_cast<int>(rank) * kTileM;
#pragma unroll 1
        for (int r = 0; r < rounds; ++r) {
          long long tw = 0;
          if (prof) tw = clock64();
          mbar_wait(empty0 + stage * 8, phase ^ 1u);
          mbar_wait(empty0 + stage * 8 + 8, phase ^ 1u);
          if (prof) t_wait += clock64() - tw;
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int dy = n_kc == 2 ? r : j, kc = n_kc == 2 ? j : 0;
              const uint32_t fb = full0 + (stage + j) * 8;
              const uint32_t a_dst = tiles_addr + (stage + j) * stage_bytes;
              const int kcol = (2 * dy * n_kc + kc) * 64;
              if (leader) mbar_arrive_expect_tx(fb, tx_bytes);
              tma_load_2d_pair(a_dst, &tmap_a, fb, kc * 64, row0 + (dy ? off_dy1 : off_dy0), kEvictNormal);
              tma_load_2d_pair(a_dst + kABytes, &tmap_b, fb, kcol, b_row0, kEvictLast);
              tma_load_2d_pair(a_dst + kABytes + b_bytes, &tmap_b, fb, kcol + n_kc * 64, b_row0, kEvictLast);
            }
          }
          __syncwarp();
          stage += 2;
          if (stage == n_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    } else
    {
      // everything loop-invariant in registers: the producer's reaction time (stage freed -> loads issued) is on the
      // critical path of the 3-stage pipeline of the wide layers
      const int off_dy0 = p.tap_off[0], off_dy1 = p.tap_off[2];
      const int n_kc = p.n_kc, kc_term = p.kc_term, lo_col = p.lo_col, dx_cols = p.n_kc * 64;
      const bool two_parts = p.n_parts == 2;
      const int b_row1 = p.n_part + b_row0;
      const uint32_t part_dst = static_cast<uint32_t>(half_rows) * 128u;
      const uint32_t n_stages = static_cast<uint32_t>(p.stages);
      uint32_t a_dst = tiles_addr;                           // shared-memory address of the current stage
      uint32_t eb = empty0, fb = full0;
      for (int tile = pair; tile < p.num_tiles; tile += n_pairs) {
        const int row0 = tile * (2 * kTileM) + static_cast<int>(rank) * kTileM;
        int kcol = 0;
#pragma unroll 1
        for (int dy = 0; dy < 2; ++dy) {
          // taps (2 dy, 2 dy + 1) read rows r and r + 1: one box of 136 rows feeds both
          const int arow = row0 + (dy ? off_dy1 : off_dy0);
          // split precision: the K chunks of a tap are [hi | hi | lo] blocks of the activation columns
          int kci = 0, term = 0;
#pragma unroll 1
          for (int kc = 0; kc < n_kc; ++kc, kcol += 64) {
            long long tw = 0;
            if (prof) tw = clock64();
            mbar_wait(eb, phase ^ 1u);
            if (prof) t_wait += clock64() - tw;
            if (elect_one()) {
              if (leader) mbar_arrive_expect_tx(fb, tx_bytes);
              tma_load_2d_pair(a_dst, &tmap_a, fb, kci * 64 + (term == 2 ? lo_col : 0), arow, kEvictNormal);
              const uint32_t b_dst = a_dst + kABytes;
              tma_load_2d_pair(b_dst, &tmap_b, fb, kcol, b_row0, kEvictLast);
              if (two_parts) tma_load_2d_pair(b_dst + part_dst, &tmap_b, fb, kcol, b_row1, kEvictLast);
              tma_load_2d_pair(b_dst + b_bytes, &tmap_b, fb, kcol + dx_cols, b_row0, kEvictLast);
              if (two_parts) tma_load_2d_pair(b_dst + b_bytes + part_dst, &tmap_b, fb, kcol + dx_cols, b_row1, kEvictLast);
            }
            __syncwarp();
            if (++kci == kc_term) {
              kci = 0;
              ++term;
            }
            a_dst += stage_bytes;
            eb += 8;
            fb += 8;
            if (++stage == n_stages) {
              stage = 0;
              phase ^= 1u;
              a_dst = tiles_addr;
              eb = empty0;
              fb = full0;
            }
          }
          kcol += dx_cols;                                   // skip the dx = 1 tap's columns of this dy
        }
      }
    }
    if (prof && lane == 0) {
      p.stats[blockIdx.x * 16 + 0] = clock64() - t_begin;   // producer total
      p.stats[blockIdx.x * 16 + 1] = t_wait;                // producer waiting for free stages
      const unsigned long long ns_end = global_ns();
      p.stats[blockIdx.x * 16 + 7] = static_cast<long long>(ns_end - ns_begin);   // producer total in ns (-> SM clock)
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA; warp-uniform loop,
    // one elected lane issues the MMAs and the commits that track them)
    if (leader) {
      const uint32_t idesc = make_idesc_16(2 * kTileM, p.n_part, 0, 0, p.ab_dtype, p.ab_dtype);
      const uint32_t part_bytes = static_cast<uint32_t>(half_rows) * 128u;
      const bool prof = p.stats != nullptr;
      // descriptor template (address field 0): low word = LBO field, high word = SBO 1024 | version 1 | SWIZZLE_128B
      const uint64_t desc0 = make_sw128_desc(0, 0, 1024);
      const uint32_t desc_lo0 = static_cast<uint32_t>(desc0), desc_hi = static_cast<uint32_t>(desc0 >> 32);
      // The dx = 1 operand starts one row (128 B) into an 8-row swizzle atom.  The hardware swizzles on absolute
      // shared-memory address bits, exactly as TMA wrote the box, so the descriptor needs no base-offset field
      // (measured on B200: base offset 1 gives wrong results, 0 is bit-exact against the oracle).
      const uint32_t desc_hi_dx1 = desc_hi;
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
      uint32_t stage = 0, phase = 0;
      int it = 0;
      long long t_full = 0, t_tmem = 0, t_issue = 0, t_commit = 0, t_begin = prof ? clock64() : 0;
      if (p.pair_issue) {
        // ---- narrow layers: two stages (both chunks of a dy, or both dy of a one-chunk tile) per round; everything
        // loop-invariant is hoisted so that a round costs ~100 instructions of bookkeeping for up to 16 MMAs
        const int k0 = p.n_kc == 2 ? 4 : p.last_ksteps, k1 = p.last_ksteps;
        const int rounds = p.n_kc;                              // rounds per tile: 2 (one per dy) or 1
        const uint32_t stage_lo = stage_bytes >> 4, bdx_lo = b_bytes >> 4;
        const uint32_t a_lo_base = desc_lo0 + ((tiles_addr & 0x3FFFFu) >> 4);
        const uint32_t n_stages = static_cast<uint32_t>(p.stages);
        for (int tile = pair; tile < p.num_tiles; tile += n_pairs, ++it) {
          const int par = quad ? (it & 3) : (it & 1);
          long long tw = 0;
          if (prof) tw = clock64();
          mbar_wait(smem_u32(&tmem_empty_bar[par]), ((quad ? it >> 2 : it >> 1) & 1) ^ 1u);
          if (prof) t_tmem += clock64() - tw;
          const uint32_t acc_base = tmem_base + (quad ? par * 128 : (par ? base1 : 0));
          const uint32_t tfull = smem_u32(&tmem_full_bar[par]);
#pragma unroll 1
          for (int r = 0; r < rounds; ++r) {
            if (prof) tw = clock64();
            mbar_wait(full0 + stage * 8, phase);
            mbar_wait(full0 + stage * 8 + 8, phase);
            if (prof) t_full += clock64() - tw;
            tc_fence_after();
            const uint32_t a0 = a_lo_base + stage * stage_lo, b0 = a0 + (kABytes >> 4);
            const uint32_t a1 = a0 + stage_lo, b1 = b0 + stage_lo;
            if (elect_one()) {
              const long long ti0 = prof ? clock64() : 0;
              umma_f16_pair_entry<1, 2>(acc_base, acc_base, a0, b0, b0, desc_hi, desc_hi, idesc, idesc, r ? 1u : 0u, k0);
              umma_f16_pair_entry<1, 2>(acc_base, acc_base, a0 + 8, b0 + bdx_lo, b0, desc_hi, desc_hi, idesc, idesc, 1u, k0);
              umma_f16_pair_entry<1, 2>(acc_base, acc_base, a1, b1, b1, desc_hi, desc_hi, idesc, idesc, 1u, k1);
              umma_f16_pair_entry<1, 2>(acc_base, acc_base, a1 + 8, b1 + bdx_lo, b1, desc_hi, desc_hi, idesc, idesc, 1u, k1);
              const long long ti1 = prof ? clock64() : 0;
              umma_commit_pair(empty0 + stage * 8);
              umma_commit_pair(empty0 + stage * 8 + 8);
              if (r == rounds - 1) umma_commit_pair(tfull);
              if (prof) {
                t_issue += ti1 - ti0;
                t_commit += clock64() - ti1;
              }
            }
            __syncwarp();
            stage += 2;
            if (stage == n_stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      } else {
      const int n_kc = p.n_kc, kc_term = p.kc_term, last_ksteps = p.last_ksteps;
      for (int tile = pair; tile < p.num_tiles; tile += n_pairs, ++it) {
        const int par = quad ? (it & 3) : (it & 1);
        long long tw = 0;
        if (prof) tw = clock64();
        mbar_wait(smem_u32(&tmem_empty_bar[par]), ((quad ? it >> 2 : it >> 1) & 1) ^ 1u);   // region drained (tile it - 2 | 4)
        // regions overlap only when n_pad > 256: then the columns shared with tile it - 1 must have been drained
        if (p.n_pad > 256 && it > 0) mbar_wait(smem_u32(tmem_ovl_bar), (it - 1) & 1);
        if (prof) t_tmem += clock64() - tw;
        tc_fence_after();
        const uint32_t acc_base = tmem_base + (quad ? par * 128 : (par ? base1 : 0));
        uint32_t accumulate = 0;
#pragma unroll 1
        for (int dy = 0; dy < 2; ++dy) {
          int kci = 0;                                         // chunk index inside the current split term
#pragma unroll 1
          for (int kc = 0; kc < n_kc; ++kc) {
            if (prof) tw = clock64();
            mbar_wait(full0 + stage * 8, phase);
            if (prof) t_full += clock64() - tw;
            tc_fence_after();
            const uint32_t a_addr = tiles_addr + stage * stage_bytes;
            int ksteps = 4;
            if (++kci == kc_term) {
              kci = 0;
              ksteps = last_ksteps;
            }
            const uint32_t a_lo = desc_lo0 + ((a_addr & 0x3FFFFu) >> 4);
            const uint32_t b_lo = desc_lo0 + (((a_addr + kABytes) & 0x3FFFFu) >> 4);
            const bool last = (dy == 1 && kc == n_kc - 1);
            if (elect_one()) {
              const long long ti0 = prof ? clock64() : 0;
              // dx = 0: rows [0, 128) of the box; dx = 1: rows [1, 129) = start address + 128 B (+8 in the descriptor)
              if (p.n_parts == 2) {
                umma_f16_pair_entry<2, 2>(acc_base, acc_base + p.n_part, a_lo, b_lo, b_lo + (part_bytes >> 4), desc_hi, desc_hi,
                                          idesc, idesc, accumulate, ksteps);
                umma_f16_pair_entry<2, 2>(acc_base, acc_base + p.n_part, a_lo + 8, b_lo + (b_bytes >> 4),
                                          b_lo + ((b_bytes + part_bytes) >> 4), desc_hi_dx1, desc_hi, idesc, idesc, 1u, ksteps);
              } else {
                umma_f16_pair_entry<1, 2>(acc_base, acc_base, a_lo, b_lo, b_lo, desc_hi, desc_hi, idesc, idesc, accumulate,
                                          ksteps);
                umma_f16_pair_entry<1, 2>(acc_base, acc_base, a_lo + 8, b_lo + (b_bytes >> 4), b_lo, desc_hi_dx1, desc_hi,
                                          idesc, idesc, 1u, ksteps);
              }
              const long long ti1 = prof ? clock64() : 0;
              umma_commit_pair(empty0 + stage * 8);
              if (last) umma_commit_pair(smem_u32(&tmem_full_bar[par]));
              if (prof) {
                t_issue += ti1 - ti0;
                t_commit += clock64() - ti1;
              }
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
      }
      if (prof && lane == 0) {
        p.stats[blockIdx.x * 16 + 2] = clock64() - t_begin;   // MMA issuer total
        p.stats[blockIdx.x * 16 + 3] = t_full;                // ... waiting for TMA data
        p.stats[blockIdx.x * 16 + 4] = t_tmem;                // ... waiting for the epilogue to free TMEM
      }
      if (prof && t_issue) {                                  // the elected lane
        p.stats[blockIdx.x * 16 + 12] = t_issue;              // ... inside the MMA asm blocks
        p.stats[blockIdx.x * 16 + 13] = t_commit;             // ... inside the commits
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;                              // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;                    // 0: even segments of a tile, 1: odd segments
    EpiWarp w;
    w.stg_addr = tiles_addr + p.epi_off + (warp - 2) * (FL::dual(p) ? 2u : 1u) * kSegBytes;
    w.zstg_addr = tiles_addr + p.epi_off + kEpiWarps * (FL::dual(p) ? 2u : 1u) * kSegBytes + (warp - 2) * kSegBytes;
    w.s_stat = (FL::stats(p) && p.stat_off) ? reinterpret_cast<float*>(smem + p.stat_off) + (warp - 2) * 2 * p.n_pad : nullptr;
    w.s_add = s_add;
    w.s_mul = s_mul;
    int it = 0;
    long long t_wait = 0, t_begin = clock64();
    const int seg_first = quad ? 0 : half, seg_step = quad ? 1 : 2;
    for (int tile = pair; tile < p.num_tiles; tile += n_pairs, ++it) {
      if (quad && (it & 1) != half) continue;             // the other warp of this lane quadrant drains this tile
      const int par = quad ? (it & 3) : (it & 1);
      const int row0 = tile * (2 * kTileM) + static_cast<int>(rank) * kTileM + q * 32;
      const int64_t s = static_cast<int64_t>(row0) + lane;
      const bool in_range = s < p.n_slots;
      const int rows_valid = p.n_slots - row0 >= 32 ? 32 : static_cast<int>(p.n_slots - row0);   // may be <= 0
      int b = 0, sy = 0, sx = 0;
      if (in_range) slot_coords(p, s, b, sy, sx);
      const bool valid = in_range && (p.type == 0 || (sy >= 1 && sx >= 1));
      // Column order: the columns shared with the other TMEM region first (the last `ovl` columns of an even tile, the
      // first ones of an odd tile).  Range 1 = [start1, n_pad), then range 2 = [0, start1); 32-column segments, this
      // warp takes the positions i = half, half + 2, ...
      const int start1 = (par == 0 && ovl > 0) ? p.n_pad - ovl : 0;
      const int nseg1 = (p.n_pad - start1 + 31) >> 5;
      const int nseg = nseg1 + ((start1 + 31) >> 5);
      const int ovl_pos = ovl > 0 ? (ovl >> 5) : nseg;          // positions after which the shared columns are free
      auto seg_c0 = [&](int i) { return i < nseg1 ? start1 + 32 * i : 32 * (i - nseg1); };
      auto seg_n = [&](int i) { const int end = i < nseg1 ? p.n_pad : start1; const int c = seg_c0(i); return end - c < 32 ? end - c : 32; };
      // last position of this warp inside the shared columns / overall (-1: none)
      const int last_ovl_i = ovl_pos > seg_first ? seg_first + (ovl_pos - 1 - seg_first) / seg_step * seg_step : -1;
      const int last_i = nseg > seg_first ? seg_first + (nseg - 1 - seg_first) / seg_step * seg_step : -1;
      // this slot's saved ReLU bits of the first two segments, fetched before the accumulators are ready (one word per
      // 32 channels); the words of the following segments are prefetched one loop iteration ahead
      const uint32_t* gate_row = (FL::gate(p) && in_range) ? p.gate_bits + s * p.ld_bits : nullptr;
      auto gate_word = [&](int i) { return (gate_row && i < nseg) ? __ldg(gate_row + (seg_c0(i) >> 5)) : 0u; };
      uint32_t g0 = gate_word(seg_first), g1 = gate_word(seg_first + seg_step);
      if (FL::bits(p) && (quad || half == 0) && in_range)                 // words beyond the channels (ld_bits > ceil(n_pad / 32))
        for (int ww = (p.n_pad + 31) >> 5; ww < p.ld_bits; ++ww) p.relu_bits[s * p.ld_bits + ww] = 0u;
      long long tw = 0;
      if (p.stats) tw = clock64();
      mbar_wait(smem_u32(&tmem_full_bar[par]), (quad ? it >> 2 : it >> 1) & 1);
      if (p.stats) t_wait += clock64() - tw;
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (quad ? par * 128 : (par ? base1 : 0));
      auto load_seg = [&](int i, uint32_t (&r)[32]) {
        if (seg_n(i) == 32) tmem_ld32(taddr + seg_c0(i), r);
        else tmem_ld16_lo(taddr + seg_c0(i), r);
      };
      auto release = [&](bool rel_ovl, bool rel_all) {
        if (rel_ovl || rel_all) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (rel_ovl) mbar_arrive_leader(smem_u32(tmem_ovl_bar));
            if (rel_all) mbar_arrive_leader(smem_u32(&tmem_empty_bar[par]));
          }
        }
      };
      release(last_ovl_i < 0, last_i < 0);               // nothing of mine in the shared columns / in this tile
      auto process = [&](const uint32_t (&r)[32], int i, uint32_t gword) {
        const int c0 = seg_c0(i), n = seg_n(i);
        if (!FL::direct(p)) epi_segment<F>(p, w, r, c0, n, gword, valid, row0, rows_valid, lane);
        else epi_box_direct(p, s_add, s_mul, r, c0, n, s, in_range, valid, b, sy, sx);
      };
      uint32_t ra[32], rb[32];
      if (seg_first < nseg) load_seg(seg_first, ra);
      // NOT unrolled: the body (two segments) is ~1000 instructions; unrolling it five times pushed the kernel far
      // beyond the instruction cache and starved the producer / MMA warps ('no_inst' stalls)
#pragma unroll 1
      for (int i = seg_first; i < nseg; i += 2 * seg_step) {
        const int i1 = i + seg_step, i2 = i + 2 * seg_step;
        const uint32_t g2 = gate_word(i2), g3 = gate_word(i2 + seg_step);
        tmem_ld_wait();
        if (i1 < nseg) load_seg(i1, rb);                     // in flight while segment i is processed
        release(i == last_ovl_i, i == last_i);
        process(ra, i, g0);
        if (i1 < nseg) {
          tmem_ld_wait();
          if (i2 < nseg) load_seg(i2, ra);
          release(i1 == last_ovl_i, i1 == last_i);
          process(rb, i1, g1);
        }
        g0 = g2;
        g1 = g3;
      }
    }
    if (p.stats && warp == 2 && lane == 0) {
      p.stats[blockIdx.x * 16 + 5] = clock64() - t_begin;     // epilogue total
      p.stats[blockIdx.x * 16 + 6] = t_wait;                  // ... waiting for accumulators
    }
  }

  if (p.stats && warp == 2 && lane == 0) p.stats[blockIdx.x * 16 + 10] = static_cast<long long>(global_ns());   // epilogue done
  tc_fence_before();
  __syncthreads();
  if (p.stat_off && p.col_sums) {
    // combine the epilogue warps' partial sums and add them to the global fp64 accumulators
    const float* st = reinterpret_cast<const float*>(smem + p.stat_off);
    for (int i = threadIdx.x; i < 2 * p.n_pad; i += kConvThreads) {
      float v = 0.f;
#pragma unroll
      for (int wq = 0; wq < kEpiWarps; ++wq) v += st[wq * 2 * p.n_pad + i];
      atomicAdd(p.col_sums + i, static_cast<double>(v));
    }
  }
  cluster_sync();                                      // nobody leaves while the peer may still touch its memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
  if (p.stats && threadIdx.x == 0) p.stats[blockIdx.x * 16 + 11] = static_cast<long long>(global_ns());          // kernel exit
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
  MMLF_REQUIRE(!(a->gate_bits || a->relu_bits) || a->ld_bits <= 16, "conv2x2: ld_bits %d too large", a->ld_bits);
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
  p.kc_term = ceil_div(a->cin_pad, 64);
  p.last_ksteps = (a->cin_pad - (p.kc_term - 1) * 64) / 16;
  MMLF_REQUIRE(a->split_in >= 0 && a->split_out >= 0, "conv2x2: negative split offset");
  if (a->split_in) {
    MMLF_REQUIRE(a->ab_dtype == kFP16, "conv2x2: split-precision operands are fp16");
    MMLF_REQUIRE(a->split_in % 8 == 0 && a->split_in >= a->cin_pad && a->ld_in >= a->split_in + a->cin_pad,
                 "conv2x2: split_in %d does not fit cin_pad %d / ld_in %d", a->split_in, a->cin_pad, a->ld_in);
  }
  if (a->split_out) {
    MMLF_REQUIRE(a->out_mode == 0 && !a->out2 && a->out_dtype == kFP16, "conv2x2: split_out needs a single fp16 out_mode-0 output");
    MMLF_REQUIRE(a->split_out % 8 == 0 && a->split_out >= a->n_pad && a->ld_out >= a->split_out + a->n_pad,
                 "conv2x2: split_out %d does not fit n_pad %d / ld_out %d", a->split_out, a->n_pad, a->ld_out);
  }
  p.n_kc = a->split_in ? 3 * p.kc_term : p.kc_term;
  p.lo_col = a->split_in;
  // Narrow layers (the 70-channel in-nets, the heads): a pipeline stage holds only 2-8 short MMAs and the issuing warp's
  // per-stage bookkeeping (~500 cycles measured against ~200 tensor cycles) bounds the kernel, so it issues two stages
  // per round: both chunks of a dy (n_kc = 2) or both dy of a tile (n_kc = 1).  Needs an even stage count.
  static const bool no_pair_issue = getenv("MMLF_NO_PAIR_ISSUE") != nullptr;    // debugging switch: one stage per round
  p.pair_issue = (p.n_parts == 1 && p.n_kc <= 2 && !a->split_in && !no_pair_issue) ? 1 : 0;
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
  p.dual = a->out2 != nullptr || a->split_out != 0;
  p.split_out = a->split_out != 0;
  p.ld_bits = a->ld_bits;
  p.bias = a->bias;
  p.scale = a->scale;
  p.shift = a->shift;
  p.gate_bits = a->gate_bits;
  p.relu_bits = a->relu_bits;
  p.col_sums = a->col_sums;
  p.bn_z = reinterpret_cast<const uint16_t*>(a->bn_z);
  p.ld_z = a->ld_z;
  p.z_dtype = a->bn_z_dtype;
  p.bn_scale = a->bn_scale;
  p.bn_shift = a->bn_shift;
  p.bn_mean = a->bn_mean;
  if (a->bn_z) {
    MMLF_REQUIRE(a->col_sums && a->out_mode == 0, "conv2x2: bn_z needs col_sums and out_mode 0");
    MMLF_REQUIRE(a->bn_scale && a->bn_shift && a->bn_mean, "conv2x2: bn_z needs bn_scale / bn_shift / bn_mean");
    MMLF_REQUIRE(a->ld_z >= a->n_pad && (a->bn_z_dtype >> 1) == 0, "conv2x2: bad ld_z %d / bn_z_dtype", a->ld_z);
  }
  p.out = a->out;
  p.out2 = a->out2;
  p.ld_out2 = a->ld_out2;
  if (a->split_out) {                                       // the residual goes through the second-output path
    p.out2 = reinterpret_cast<uint16_t*>(a->out) + a->split_out;
    p.ld_out2 = a->ld_out;
    p.out2_dtype = kFP16;
  }
  p.stats = nullptr;
  p.stages = 0;
  p.epi_off = p.stat_off = p.aux_off = 0;
  return 0;
}

}  // namespace mmlf

using namespace mmlf;

static long long* g_conv_stats = nullptr;
// debug hook (not part of the public header): per-CTA cycle counters of the next conv launches, [grid][16] int64
extern "C" void mmlf_debug_conv_stats(long long* device_buf) { g_conv_stats = device_buf; }

typedef void (*ConvKernel)(CUtensorMap, CUtensorMap, ConvParams);
struct ConvVariant {
  uint32_t mask;
  ConvKernel fn;
};
#define MMLF_CONV_VARIANT(m) {static_cast<uint32_t>(m), conv2x2_tc2_kernel<static_cast<uint32_t>(m)>}
// the epilogue combinations the network uses (fp16 and bf16 activation storage) + the generic instance (last)
static const ConvVariant kConvVariants[] = {
    MMLF_CONV_VARIANT(kFRelu | kFOutF16),                          // eval / no-grad: first conv of a block
    MMLF_CONV_VARIANT(kFScale | kFRelu | kFOutF16),                // eval: second conv with folded BatchNorm
    MMLF_CONV_VARIANT(kFRelu | kFOutF16 | kFDual | kFOut2F16 | kFSplit),            // split-precision eval (hi + lo outputs)
    MMLF_CONV_VARIANT(kFScale | kFRelu | kFOutF16 | kFDual | kFOut2F16 | kFSplit),
    MMLF_CONV_VARIANT(kFRelu | kFBits | kFDual | kFOutF16),        // training: first conv (fp16 + bf16 copy + bits)
    MMLF_CONV_VARIANT(kFStats | kFOutF16),                         // training: second conv with batch statistics
    MMLF_CONV_VARIANT(kFRelu),
    MMLF_CONV_VARIANT(kFScale | kFRelu),
    MMLF_CONV_VARIANT(kFRelu | kFBits),
    MMLF_CONV_VARIANT(kFStats),
    MMLF_CONV_VARIANT(0),                                          // data gradient
    MMLF_CONV_VARIANT(kFGate | kFStats),                           // data gradient through a ReLU + bias gradient
    MMLF_CONV_VARIANT(kFStats | kFBnBwd),                          // data gradient + BatchNorm-backward statistics
    MMLF_CONV_VARIANT(kFDirect),                                   // heads (fp32 outputs)
    MMLF_CONV_VARIANT(kFGeneric),
};
constexpr int kNumConvVariants = sizeof(kConvVariants) / sizeof(kConvVariants[0]);

extern "C" int mmlf_conv2x2(const mmlf_conv_args* a, void* stream) {
  ConvParams p;
  if (int rc = fill_params(a, p)) return rc;
  p.stats = g_conv_stats;
  const uint32_t stage_bytes = kABytes + 2u * p.n_pad * 64u;
  // shared-memory plan behind the operand stages: [epilogue staging | statistics | barriers + constants]
  const uint32_t epi_bytes = p.out_mode == 0 ? kEpiWarps * ((p.dual ? 2u : 1u) + (p.bn_z ? 1u : 0u)) * kSegBytes : 0u;
  const uint32_t stat_bytes = p.col_sums ? kEpiWarps * 2u * p.n_pad * 4u : 0u;   // per epilogue warp: [2][n_pad] f32
  const uint32_t aux_bytes = 256 + 2 * kMaxNPad * 4 + 64;   // barriers + TMEM pointer, then the per-channel constants
  const uint32_t tail_bytes = epi_bytes + ((stat_bytes + 15u) & ~15u) + aux_bytes;
  const uint32_t max_smem = 232448;   // 227 KB opt-in limit per CTA on sm_100
  int stages = static_cast<int>((max_smem - 1024 - tail_bytes) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (p.pair_issue) stages &= ~1;                         // rounds of two stages
  MMLF_REQUIRE(stages >= 2, "conv2x2: not enough shared memory for a 2-stage pipeline (n_pad %d)", p.n_pad);
  p.stages = stages;
  p.epi_off = stages * stage_bytes;                       // multiple of 1024 (stage_bytes = 17408 + n_pad * 128)
  p.stat_off = stat_bytes ? p.epi_off + epi_bytes : 0;
  p.aux_off = p.epi_off + epi_bytes + ((stat_bytes + 15u) & ~15u);
  const uint32_t smem_bytes = 1024 + p.aux_off + aux_bytes;

  CUtensorMap tmap_a, tmap_b;
  // (split precision: the map also covers the lo block; chunk columns past cin_pad are never consumed by a k-step)
  if (int rc = make_tmap_2d_16(&tmap_a, a->in, a->split_in ? a->split_in + a->cin_pad : a->cin_pad, p.n_slots,
                               static_cast<uint64_t>(a->ld_in) * 2, 64, kABoxRows, 128))
    return rc;
  const uint64_t k_total = static_cast<uint64_t>(4) * p.n_kc * 64;
  if (int rc = make_tmap_2d_16(&tmap_b, a->wpack, k_total, p.n_pad, k_total * 2, 64, p.n_part / 2, 128)) return rc;

  static bool attr_set = false;
  if (!attr_set) {
    for (int i = 0; i < kNumConvVariants; ++i) {
      cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void*>(kConvVariants[i].fn),
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
      MMLF_REQUIRE(e == cudaSuccess, "conv2x2: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    attr_set = true;
  }
  uint32_t mask;
  if (p.out_mode != 0) {
    mask = kFDirect;
  } else {
    mask = (p.has_scale ? kFScale : 0u) | (p.relu ? kFRelu : 0u) | (p.gate_bits ? kFGate : 0u) |
           (p.relu_bits ? kFBits : 0u) | (p.dual ? kFDual : 0u) | (p.col_sums ? kFStats : 0u) |
           (p.out_dtype == kFP16 ? kFOutF16 : 0u) | (p.dual && p.out2_dtype == kFP16 ? kFOut2F16 : 0u) |
           (p.split_out ? kFSplit : 0u) | (p.bn_z ? kFBnBwd : 0u);
  }
  ConvKernel fn = kConvVariants[kNumConvVariants - 1].fn;
  for (int i = 0; i < kNumConvVariants - 1; ++i)
    if (kConvVariants[i].mask == mask) fn = kConvVariants[i].fn;
  const int max_pairs = sm_count() / 2;
  const int pairs = p.num_tiles < max_pairs ? p.num_tiles : max_pairs;
  void* args[] = {&tmap_a, &tmap_b, &p};
  cudaError_t e = cudaLaunchKernel(reinterpret_cast<const void*>(fn), dim3(2 * pairs), dim3(kConvThreads), args, smem_bytes,
                                   static_cast<cudaStream_t>(stream));
  MMLF_REQUIRE(e == cudaSuccess, "conv2x2_tc2_kernel: %s", cudaGetErrorString(e));
  return check_launch("conv2x2_tc2_kernel");
}

extern "C" int mmlf_conv2x2_simt(const mmlf_conv_args* a, void* stream) {
  ConvParams p;
  if (int rc = fill_params(a, p)) return rc;
  MMLF_REQUIRE(!(a->out2 || a->relu_bits || a->col_sums), "conv2x2_simt: out2 / relu_bits / col_sums are not supported");
  MMLF_REQUIRE(!(a->split_in || a->split_out || a->bn_z), "conv2x2_simt: split precision / BN statistics are not supported");
  const int64_t total = p.n_slots * (p.n_pad / 16);
  const int threads = 128;
  const int64_t blocks = ceil_div64(total, threads);
  conv2x2_simt_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(a->in), a->ld_in, a->cin_pad,
      reinterpret_cast<const __nv_bfloat16*>(a->wpack), 4 * p.n_kc * 64, p);
  return check_launch("conv2x2_simt_kernel");
}
