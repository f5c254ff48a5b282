// Host-side helpers shared by the C-ABI translation units: error reporting, launch checks,
// and TMA tensor-map creation through the driver entry point (no link-time libcuda dependency,
// so the library loads on a machine without a GPU driver).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

namespace mmlf {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

// 2-D tensor map over 16-bit elements (bf16 / fp16 move identically): dims {inner, outer}, row pitch in bytes,
// box {box_inner, box_outer}, 64 or 128 B swizzle; out-of-bounds elements read as zero and are not written.
int make_tmap_2d_16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                    uint32_t box_inner, uint32_t box_outer, int swizzle_bytes);

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
int sm_count();

#define MMLF_REQUIRE(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      ::mmlf::set_error(__VA_ARGS__);    \
      return 1;                          \
    }                                    \
  } while (0)

}  // namespace mmlf
