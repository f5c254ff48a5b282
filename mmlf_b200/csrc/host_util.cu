#include "host_util.h"

#include <string.h>

#include <mutex>

#include "../../include/mmlf_b200.h"

namespace mmlf {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return 2;
  }
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_2d_16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                    uint32_t box_inner, uint32_t box_outer, int swizzle_bytes) {
  EncodeTiledFn enc = get_encode();
  MMLF_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
  MMLF_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  MMLF_REQUIRE((pitch_bytes & 15) == 0, "TMA row pitch must be a multiple of 16 bytes (got %llu)",
               (unsigned long long)pitch_bytes);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  MMLF_REQUIRE(swizzle_bytes == 128 || swizzle_bytes == 64, "TMA swizzle must be 64 or 128 bytes");
  MMLF_REQUIRE(box_inner * 2 <= static_cast<uint32_t>(swizzle_bytes), "TMA box rows must fit the swizzle span");
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MMLF_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (inner %llu outer %llu pitch %llu box %u x %u)",
               (int)r, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_bytes,
               box_inner, box_outer);
  return 0;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

__global__ void vec_jobs_kernel(const mmlf_vec_job* __restrict__ jobs) {
  const mmlf_vec_job j = jobs[blockIdx.x];
  for (int i = threadIdx.x; i < j.n; i += blockDim.x) {
    float v = (j.src_f64 & 1) ? static_cast<float>(static_cast<const double*>(j.src)[i])
                              : static_cast<const float*>(j.src)[i];
    if (j.src2)
      v += (j.src_f64 & 2) ? static_cast<float>(static_cast<const double*>(j.src2)[i])
                           : static_cast<const float*>(j.src2)[i];
    j.dst[i] = j.accumulate ? j.dst[i] + v : v;
  }
}

}  // namespace mmlf

extern "C" {

int mmlf_zero(void* ptr, int64_t bytes, void* stream) {
  MMLF_REQUIRE(ptr != nullptr && bytes >= 0, "zero: bad arguments");
  cudaError_t e = cudaMemsetAsync(ptr, 0, static_cast<size_t>(bytes), static_cast<cudaStream_t>(stream));
  MMLF_REQUIRE(e == cudaSuccess, "zero: %s", cudaGetErrorString(e));
  return 0;
}

int mmlf_vec_jobs(const mmlf_vec_job* jobs, int n_jobs, void* stream) {
  MMLF_REQUIRE(jobs != nullptr && n_jobs >= 0, "vec_jobs: bad arguments");
  if (n_jobs == 0) return 0;
  mmlf::vec_jobs_kernel<<<n_jobs, 128, 0, static_cast<cudaStream_t>(stream)>>>(jobs);
  return mmlf::check_launch("vec_jobs");
}

const char* mmlf_last_error(void) { return mmlf::g_err; }
int mmlf_abi_version(void) { return MMLF_ABI_VERSION; }

int mmlf_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    mmlf::set_error("cudaGetDevice: %s", cudaGetErrorString(e));
    return 2;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  MMLF_REQUIRE(major == 10, "mmlf_b200 kernels are built for sm_100a only; device is sm_%d%d", major, minor);
  return 0;
}

}  // extern "C"
