// C ABI of libemdr2_b200.so (see include/emdr2_b200.h for the contract and the reference
// interfaces each entry point replaces).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/emdr2_b200.h"
#include "mips_merge.cuh"
#include "mips_scan.cuh"
#include "ptx.cuh"
#include "capi_common.cuh"

namespace {

using emdr2::capi::fail;
using emdr2::capi::make_tmap_2d;
using emdr2::capi::DeviceGuard;

struct MipsHandle {
  uint32_t magic = 0x4d495053u;  // "MIPS"
  int d = 0, dtype = 0, device = 0;
  int sm_count = 0, max_smem = 0;
  uint32_t num_kb = 0, num_stages = 0, smem_bytes = 0;

  const void* rows = nullptr;
  const int64_t* ids = nullptr;
  int64_t n = -1, id_base = 0;
  CUtensorMap tmap_e;

  // device workspace
  uint64_t* gmax = nullptr;
  uint64_t* gthr = nullptr;
  unsigned long long* stats = nullptr;
  float* pool_scores = nullptr;
  int64_t* pool_ids = nullptr;
  uint32_t* pool_cnt = nullptr;
  uint32_t epoch = 0;

  // staging for the host-buffer entry point
  void* stage_q = nullptr;
  float* stage_scores = nullptr;
  int64_t* stage_ids = nullptr;
  int stage_nq = 0, stage_k = 0;

  // options / last-launch facts
  int opt_probe = 1, opt_share = 1, opt_max_ctas = 0, opt_stats = 0;
  uint32_t probe_timeout_ns = 30000;
  int last_ctas = 0, last_tiles = 0;

  // optional per-launch timing of the scan kernel (option "timing"): event pairs recorded on the
  // caller's stream around every scan launch, summed by get_stat("scan_ns").
  int opt_timing = 0;
  std::vector<cudaEvent_t> ev;  // [2 * launches]
  size_t ev_used = 0;
};

MipsHandle* as_handle(void* h) {
  MipsHandle* p = static_cast<MipsHandle*>(h);
  return (p && p->magic == 0x4d495053u) ? p : nullptr;
}

void free_workspace(MipsHandle* h) {
  cudaFree(h->gmax);
  cudaFree(h->gthr);
  cudaFree(h->stats);
  cudaFree(h->pool_scores);
  cudaFree(h->pool_ids);
  cudaFree(h->pool_cnt);
  cudaFree(h->stage_q);
  cudaFree(h->stage_scores);
  cudaFree(h->stage_ids);
  for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
  h->ev.clear();
}

}  // namespace

extern "C" {

const char* emdr2_last_error(void) { return emdr2::capi::last_error(); }

const char* emdr2_version(void) {
  static char buf[96];
  snprintf(buf, sizeof(buf), "emdr2_b200 0.1.0 sm_100a nvcc %d.%d", __CUDACC_VER_MAJOR__,
           __CUDACC_VER_MINOR__);
  return buf;
}

int emdr2_mips_create(int d, int dtype, int device, void** out_handle) {
  if (!out_handle) return fail(EMDR2_EINVAL, "out_handle is NULL");
  *out_handle = nullptr;
  if (d < 8 || d > 1024 || (d % 8) != 0)
    return fail(EMDR2_EINVAL, "embedding dimension d=%d must be a multiple of 8 in [8, 1024]", d);
  if (dtype != EMDR2_DTYPE_FP16 && dtype != EMDR2_DTYPE_BF16)
    return fail(EMDR2_EINVAL, "dtype %d is not EMDR2_DTYPE_FP16/BF16", dtype);
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev)
    return fail(EMDR2_EINVAL, "device %d out of range (%d CUDA devices visible)", device, ndev);
  DeviceGuard guard(device);
  if (!guard.ok) return fail(EMDR2_ECUDA, "cudaSetDevice(%d) failed", device);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(EMDR2_EUNSUPPORTED,
                "device %d is sm_%d%d; this library contains sm_100a (B200) code only", device,
                prop.major, prop.minor);

  MipsHandle* h = new (std::nothrow) MipsHandle();
  if (!h) return fail(EMDR2_ENOMEM, "host allocation failed");
  h->d = d;
  h->dtype = dtype;
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->max_smem = static_cast<int>(prop.sharedMemPerBlockOptin);
  h->num_kb = static_cast<uint32_t>((d + emdr2::kBlockK - 1) / emdr2::kBlockK);

  // pipeline depth from whatever shared memory the resident queries and candidate lists leave
  int stages = emdr2::kMaxStages;
  while (stages > 0 &&
         emdr2::scan_smem_layout(h->num_kb, static_cast<uint32_t>(stages)).total >
             static_cast<uint32_t>(h->max_smem))
    --stages;
  if (stages < 2) {
    delete h;
    return fail(EMDR2_EUNSUPPORTED, "d=%d leaves room for %d pipeline stages (<2)", d, stages);
  }
  h->num_stages = static_cast<uint32_t>(stages);
  h->smem_bytes = emdr2::scan_smem_layout(h->num_kb, h->num_stages).total;

  const size_t nparts = static_cast<size_t>(emdr2::kMaxCtas) * emdr2::kQ;
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = cudaMalloc(&h->gmax, nparts * sizeof(uint64_t));
  if (e == cudaSuccess) e = cudaMalloc(&h->gthr, emdr2::kQ * sizeof(uint64_t));
  if (e == cudaSuccess) e = cudaMalloc(&h->stats, 4 * sizeof(unsigned long long));
  const size_t npool = static_cast<size_t>(emdr2::kQ) * emdr2::kPoolCap;
  if (e == cudaSuccess) e = cudaMalloc(&h->pool_scores, npool * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&h->pool_ids, npool * sizeof(int64_t));
  if (e == cudaSuccess) e = cudaMalloc(&h->pool_cnt, emdr2::kQ * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemset(h->pool_cnt, 0, emdr2::kQ * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemset(h->gmax, 0, nparts * sizeof(uint64_t));
  if (e == cudaSuccess) e = cudaMemset(h->gthr, 0, emdr2::kQ * sizeof(uint64_t));
  if (e == cudaSuccess) e = cudaMemset(h->stats, 0, 4 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = emdr2::mips_scan_prepare(h->smem_bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();  // one-time: workspace zeroing visible
  if (e != cudaSuccess) {
    free_workspace(h);
    delete h;
    return fail(e == cudaErrorMemoryAllocation ? EMDR2_ENOMEM : EMDR2_ECUDA,
                "workspace setup failed: %s", cudaGetErrorString(e));
  }
  *out_handle = h;
  return EMDR2_OK;
}

int emdr2_mips_destroy(void* handle) {
  MipsHandle* h = as_handle(handle);
  if (!h) return fail(EMDR2_EINVAL, "invalid handle");
  DeviceGuard guard(h->device);
  free_workspace(h);
  h->magic = 0;
  delete h;
  return EMDR2_OK;
}

int emdr2_mips_set_shard(void* handle, const void* dev_rows, const int64_t* dev_ids, int64_t n,
                         int64_t id_base) {
  MipsHandle* h = as_handle(handle);
  if (!h) return fail(EMDR2_EINVAL, "invalid handle");
  if (n < 0) return fail(EMDR2_EINVAL, "n=%lld is negative", static_cast<long long>(n));
  if (n > 0x7fffff00ll)
    return fail(EMDR2_EINVAL, "n=%lld rows exceed the per-shard limit of 2^31-256",
                static_cast<long long>(n));
  if (n > 0) {
    if (!dev_rows) return fail(EMDR2_EINVAL, "dev_rows is NULL with n=%lld", static_cast<long long>(n));
    if (reinterpret_cast<uintptr_t>(dev_rows) % 16 != 0)
      return fail(EMDR2_EINVAL, "dev_rows must be 16-byte aligned");
    int rc = make_tmap_2d(&h->tmap_e, h->dtype, dev_rows, static_cast<uint64_t>(n),
                          static_cast<uint64_t>(h->d), static_cast<uint64_t>(h->d), emdr2::kTileN);
    if (rc != EMDR2_OK) return rc;
  }
  h->rows = dev_rows;
  h->ids = dev_ids;
  h->n = n;
  h->id_base = id_base;
  return EMDR2_OK;
}

int emdr2_mips_search(void* handle, const void* dev_q, int nq, int k, float* dev_scores,
                      int64_t* dev_ids, void* cuda_stream) {
  MipsHandle* h = as_handle(handle);
  if (!h) return fail(EMDR2_EINVAL, "invalid handle");
  if (h->n < 0) return fail(EMDR2_ESTATE, "emdr2_mips_search called before emdr2_mips_set_shard");
  if (nq < 0) return fail(EMDR2_EINVAL, "nq=%d is negative", nq);
  if (k < 1 || k > EMDR2_MIPS_MAX_K)
    return fail(EMDR2_EINVAL, "k=%d outside [1, %d]", k, EMDR2_MIPS_MAX_K);
  if (nq == 0) return EMDR2_OK;
  if (!dev_q || !dev_scores || !dev_ids)
    return fail(EMDR2_EINVAL, "NULL query/output pointer");
  if (reinterpret_cast<uintptr_t>(dev_q) % 16 != 0)
    return fail(EMDR2_EINVAL, "dev_q must be 16-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  DeviceGuard guard(h->device);
  if (!guard.ok) return fail(EMDR2_ECUDA, "cudaSetDevice(%d) failed", h->device);

  if (h->n == 0) {
    CUDA_TRY(emdr2::launch_mips_fill_empty(dev_scores, dev_ids, nq * k, stream));
    return EMDR2_OK;
  }

  const uint32_t num_tiles = static_cast<uint32_t>((h->n + emdr2::kTileN - 1) / emdr2::kTileN);
  int grid = h->sm_count;
  if (h->opt_max_ctas > 0 && h->opt_max_ctas < grid) grid = h->opt_max_ctas;
  if (grid > emdr2::kMaxCtas) grid = emdr2::kMaxCtas;
  if (static_cast<uint32_t>(grid) > num_tiles) grid = static_cast<int>(num_tiles);
  h->last_ctas = grid;
  h->last_tiles = static_cast<int>(num_tiles);

  const size_t elt = 2;
  for (int q0 = 0; q0 < nq; q0 += emdr2::kQ) {
    const int nq_pass = (nq - q0) < emdr2::kQ ? (nq - q0) : emdr2::kQ;
    const uint8_t* qptr = static_cast<const uint8_t*>(dev_q) + static_cast<size_t>(q0) * h->d * elt;
    CUtensorMap tmap_q;
    int rc = make_tmap_2d(&tmap_q, h->dtype, qptr, static_cast<uint64_t>(nq_pass),
                          static_cast<uint64_t>(h->d), static_cast<uint64_t>(h->d), emdr2::kQ);
    if (rc != EMDR2_OK) return rc;

    emdr2::ScanArgs a;
    memset(&a, 0, sizeof(a));
    a.n_rows = static_cast<uint32_t>(h->n);
    a.nq = static_cast<uint32_t>(nq_pass);
    a.k = static_cast<uint32_t>(k);
    a.num_kb = h->num_kb;
    a.num_stages = h->num_stages;
    a.num_tiles = num_tiles;
    a.idesc = emdr2::ptx::instr_desc_f16(h->dtype == EMDR2_DTYPE_BF16 ? 1 : 0, emdr2::kQ,
                                         emdr2::kTileN);
    a.epoch = ++h->epoch;
    if (h->epoch == 0xffffffffu) h->epoch = 0;  // 0 is "never written"; wrap far before overflow
    const bool share = h->opt_share && grid >= k;
    a.flags = (share ? emdr2::kFlagShare : 0u) |
              ((share && h->opt_probe && num_tiles >= 2u * static_cast<uint32_t>(grid))
                   ? emdr2::kFlagProbe
                   : 0u);
    a.probe_timeout_ns = h->probe_timeout_ns;
    a.ids = h->ids;
    a.id_base = h->id_base;
    a.pool_scores = h->pool_scores;
    a.pool_ids = h->pool_ids;
    a.pool_cnt = h->pool_cnt;
    a.pool_cap = emdr2::kPoolCap;
    a.gmax = h->gmax;
    a.gthr = h->gthr;
    a.stats = h->opt_stats ? h->stats : nullptr;
    if (h->opt_stats) CUDA_TRY(cudaMemsetAsync(h->stats, 0, 4 * sizeof(unsigned long long), stream));

    if (h->opt_timing) {
      if (h->ev_used + 2 > h->ev.size()) {
        cudaEvent_t e0, e1;
        CUDA_TRY(cudaEventCreate(&e0));
        CUDA_TRY(cudaEventCreate(&e1));
        h->ev.push_back(e0);
        h->ev.push_back(e1);
      }
      CUDA_TRY(cudaEventRecord(h->ev[h->ev_used], stream));
    }
    emdr2::launch_mips_scan(tmap_q, h->tmap_e, a, grid, h->smem_bytes, stream);
    CUDA_TRY(cudaGetLastError());
    if (h->opt_timing) {
      CUDA_TRY(cudaEventRecord(h->ev[h->ev_used + 1], stream));
      h->ev_used += 2;
    }
    CUDA_TRY(emdr2::launch_mips_merge_pool(h->pool_scores, h->pool_ids, h->pool_cnt,
                                           emdr2::kPoolCap, nq_pass, k,
                                           dev_scores + static_cast<size_t>(q0) * k,
                                           dev_ids + static_cast<size_t>(q0) * k, stream));
  }
  return EMDR2_OK;
}

int emdr2_mips_search_host(void* handle, const void* host_q, int nq, int k, float* host_scores,
                           int64_t* host_ids, void* cuda_stream) {
  MipsHandle* h = as_handle(handle);
  if (!h) return fail(EMDR2_EINVAL, "invalid handle");
  if (nq < 0) return fail(EMDR2_EINVAL, "nq=%d is negative", nq);
  if (k < 1 || k > EMDR2_MIPS_MAX_K)
    return fail(EMDR2_EINVAL, "k=%d outside [1, %d]", k, EMDR2_MIPS_MAX_K);
  if (nq == 0) return EMDR2_OK;
  if (!host_q || !host_scores || !host_ids) return fail(EMDR2_EINVAL, "NULL host pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  DeviceGuard guard(h->device);
  if (!guard.ok) return fail(EMDR2_ECUDA, "cudaSetDevice(%d) failed", h->device);
  if (nq > h->stage_nq || k > h->stage_k) {
    const int cap_nq = nq > h->stage_nq ? nq : h->stage_nq;
    const int cap_k = k > h->stage_k ? k : h->stage_k;
    CUDA_TRY(cudaStreamSynchronize(stream));
    cudaFree(h->stage_q);
    cudaFree(h->stage_scores);
    cudaFree(h->stage_ids);
    h->stage_q = nullptr;
    h->stage_scores = nullptr;
    h->stage_ids = nullptr;
    h->stage_nq = h->stage_k = 0;
    CUDA_TRY(cudaMalloc(&h->stage_q, static_cast<size_t>(cap_nq) * h->d * 2));
    CUDA_TRY(cudaMalloc(&h->stage_scores, static_cast<size_t>(cap_nq) * cap_k * sizeof(float)));
    CUDA_TRY(cudaMalloc(&h->stage_ids, static_cast<size_t>(cap_nq) * cap_k * sizeof(int64_t)));
    h->stage_nq = cap_nq;
    h->stage_k = cap_k;
  }
  CUDA_TRY(cudaMemcpyAsync(h->stage_q, host_q, static_cast<size_t>(nq) * h->d * 2,
                           cudaMemcpyHostToDevice, stream));
  int rc = emdr2_mips_search(handle, h->stage_q, nq, k, h->stage_scores, h->stage_ids, cuda_stream);
  if (rc != EMDR2_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(host_scores, h->stage_scores, static_cast<size_t>(nq) * k * sizeof(float),
                           cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaMemcpyAsync(host_ids, h->stage_ids, static_cast<size_t>(nq) * k * sizeof(int64_t),
                           cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  return EMDR2_OK;
}

int emdr2_mips_merge(const float* dev_scores, const int64_t* dev_ids, int parts, int nq, int k,
                     float* dev_out_scores, int64_t* dev_out_ids, void* cuda_stream) {
  if (parts < 1 || nq < 0 || k < 1)
    return fail(EMDR2_EINVAL, "bad merge shape parts=%d nq=%d k=%d", parts, nq, k);
  if (nq == 0) return EMDR2_OK;
  if (!dev_scores || !dev_ids || !dev_out_scores || !dev_out_ids)
    return fail(EMDR2_EINVAL, "NULL pointer passed to emdr2_mips_merge");
  if (k >= emdr2::kMergeSortCap / 2)
    return fail(EMDR2_EINVAL, "k=%d too large for the merge kernel (< %d)", k,
                emdr2::kMergeSortCap / 2);
  CUDA_TRY(emdr2::launch_mips_merge_dense(dev_scores, dev_ids, parts, nq, k, dev_out_scores,
                                          dev_out_ids, static_cast<cudaStream_t>(cuda_stream)));
  return EMDR2_OK;
}

int emdr2_mips_set_option(void* handle, const char* name, int64_t value) {
  MipsHandle* h = as_handle(handle);
  if (!h || !name) return fail(EMDR2_EINVAL, "invalid handle or option name");
  if (!strcmp(name, "probe")) h->opt_probe = value != 0;
  else if (!strcmp(name, "share")) h->opt_share = value != 0;
  else if (!strcmp(name, "max_ctas")) h->opt_max_ctas = static_cast<int>(value);
  else if (!strcmp(name, "stats")) h->opt_stats = value != 0;
  else if (!strcmp(name, "probe_timeout_ns")) h->probe_timeout_ns = static_cast<uint32_t>(value);
  else if (!strcmp(name, "timing")) {
    h->opt_timing = value != 0;
    h->ev_used = 0;
  }
  else return fail(EMDR2_EINVAL, "unknown option '%s'", name);
  return EMDR2_OK;
}

int emdr2_mips_get_stat(void* handle, const char* name, int64_t* out_value) {
  MipsHandle* h = as_handle(handle);
  if (!h || !name || !out_value) return fail(EMDR2_EINVAL, "invalid handle, name or out pointer");
  if (!strcmp(name, "ctas")) *out_value = h->last_ctas;
  else if (!strcmp(name, "tiles")) *out_value = h->last_tiles;
  else if (!strcmp(name, "stages")) *out_value = h->num_stages;
  else if (!strcmp(name, "smem_bytes")) *out_value = h->smem_bytes;
  else if (!strcmp(name, "sm_count")) *out_value = h->sm_count;
  else if (!strcmp(name, "scan_launches")) *out_value = static_cast<int64_t>(h->ev_used / 2);
  else if (!strcmp(name, "scan_ns")) {
    // sum of the scan kernel's launch durations since "timing" was switched on (blocks until the
    // last recorded launch has finished), then restarts the accumulation
    DeviceGuard guard(h->device);
    double total_ms = 0.0;
    for (size_t i = 0; i + 1 < h->ev_used; i += 2) {
      float ms = 0.f;
      CUDA_TRY(cudaEventSynchronize(h->ev[i + 1]));
      CUDA_TRY(cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]));
      total_ms += ms;
    }
    h->ev_used = 0;
    *out_value = static_cast<int64_t>(total_ms * 1e6);
  }
  else if (!strcmp(name, "appends") || !strcmp(name, "compactions") ||
           !strcmp(name, "probe_wait_ns") || !strcmp(name, "probe_wait_sum_ns")) {
    DeviceGuard guard(h->device);
    unsigned long long s[4];
    CUDA_TRY(cudaMemcpy(s, h->stats, sizeof(s), cudaMemcpyDeviceToHost));
    *out_value = static_cast<int64_t>(!strcmp(name, "appends") ? s[0]
                                      : !strcmp(name, "compactions") ? s[1]
                                      : !strcmp(name, "probe_wait_ns") ? s[2] : s[3]);
  } else return fail(EMDR2_EINVAL, "unknown stat '%s'", name);
  return EMDR2_OK;
}

}  // extern "C"
