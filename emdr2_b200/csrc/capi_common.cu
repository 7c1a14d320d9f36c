#include "capi_common.cuh"

#include <cstdarg>
#include <cstdio>
#include <string>

namespace emdr2 {
namespace capi {

namespace {
thread_local std::string g_last_error;

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) !=
            cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
}  // namespace

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

const char* last_error() { return g_last_error.c_str(); }

int make_tmap_2d(CUtensorMap* out, int dtype, const void* base, uint64_t rows, uint64_t cols,
                 uint64_t ld, uint32_t box_rows, bool half_width) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(EMDR2_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {ld * 2};
  const cuuint32_t box[2] = {half_width ? 32u : 64u, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt =
      dtype == EMDR2_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = enc(out, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   half_width ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(EMDR2_ECUDA,
                "cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu ld=%llu)",
                static_cast<int>(r), static_cast<unsigned long long>(rows),
                static_cast<unsigned long long>(cols), static_cast<unsigned long long>(ld));
  return EMDR2_OK;
}

int make_tmap_3d(CUtensorMap* out, int dtype, const void* base, uint64_t batch, uint64_t rows,
                 uint64_t cols, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(EMDR2_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t gdim[3] = {cols, rows, batch};
  const cuuint64_t gstride[2] = {ld * 2, rows * ld * 2};
  const cuuint32_t box[3] = {64u, box_rows, 1u};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapDataType dt =
      dtype == EMDR2_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = enc(out, dt, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(EMDR2_ECUDA,
                "cuTensorMapEncodeTiled(3d) failed with CUresult %d (batch=%llu rows=%llu cols=%llu ld=%llu)",
                static_cast<int>(r), static_cast<unsigned long long>(batch),
                static_cast<unsigned long long>(rows), static_cast<unsigned long long>(cols),
                static_cast<unsigned long long>(ld));
  return EMDR2_OK;
}

int current_device_info(DeviceInfo* out) {
  static DeviceInfo cache[64];
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(EMDR2_EINVAL, "device ordinal %d out of range", dev);
  // Bind the device's primary context to THIS host thread: driver entry points such as
  // cuTensorMapEncodeTiled need a current context, and callers like PyTorch's autograd engine run
  // backward passes on worker threads that have only ever used the runtime API lazily.
  static thread_local int bound_dev = -1;
  if (bound_dev != dev) {
    CUDA_TRY(cudaSetDevice(dev));
    CUDA_TRY(cudaFree(nullptr));
    bound_dev = dev;
  }
  if (cache[dev].device != dev) {
    DeviceInfo d;
    CUDA_TRY(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev));
    CUDA_TRY(cudaDeviceGetAttribute(&d.major, cudaDevAttrComputeCapabilityMajor, dev));
    CUDA_TRY(cudaDeviceGetAttribute(&d.minor, cudaDevAttrComputeCapabilityMinor, dev));
    CUDA_TRY(cudaDeviceGetAttribute(&d.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    d.device = dev;
    cache[dev] = d;
  }
  *out = cache[dev];
  return EMDR2_OK;
}

}  // namespace capi
}  // namespace emdr2
