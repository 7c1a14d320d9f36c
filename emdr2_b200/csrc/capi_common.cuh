// Helpers shared by the C-ABI translation units (error reporting, TMA descriptors, device guard).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/emdr2_b200.h"

namespace emdr2 {
namespace capi {

// Records a printf-style message as the calling thread's last error and returns `code`.
int fail(int code, const char* fmt, ...);
const char* last_error();

// 2-D row-major [rows, cols] 16-bit tensor with a row pitch of `ld` elements; box = [box_rows, 64
// columns], 128-byte swizzle, out-of-bounds elements read as zero / are clipped on store.
// half_width: box = [box_rows, 32 columns] with the 64-byte swizzle instead.
int make_tmap_2d(CUtensorMap* out, int dtype, const void* base, uint64_t rows, uint64_t cols,
                 uint64_t ld, uint32_t box_rows, bool half_width = false);

// 3-D view [batch, rows, cols] of a row-major 16-bit buffer: element (b, r, c) at
// base + (b * rows + r) * ld + c.  Box = [1, box_rows, 64 columns], 128-byte swizzle.
int make_tmap_3d(CUtensorMap* out, int dtype, const void* base, uint64_t batch, uint64_t rows,
                 uint64_t cols, uint64_t ld, uint32_t box_rows);

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
    if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// Properties of the current device, cached per device ordinal.
struct DeviceInfo {
  int device = -1, sm_count = 0, major = 0, minor = 0, max_smem = 0;
};
int current_device_info(DeviceInfo* out);

}  // namespace capi
}  // namespace emdr2

#define CUDA_TRY(expr)                                                                       \
  do {                                                                                       \
    cudaError_t e_ = (expr);                                                                 \
    if (e_ != cudaSuccess)                                                                   \
      return ::emdr2::capi::fail(EMDR2_ECUDA, "%s failed: %s (%s:%d)", #expr,                \
                                 cudaGetErrorString(e_), __FILE__, __LINE__);                \
  } while (0)
