#include "host_util.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/dinounet_b200.h"
#include "gemm_common.h"

namespace b2u {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(-10, "%s: launch failed: %s", what, cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int encode_tensor_map(CUtensorMap* map, int dtype, int rank, const void* base, const cuuint64_t* dims,
                      const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* elem_strides) {
  EncodeTiledFn fn = get_encode();
  if (!fn) return set_error(-20, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  if (reinterpret_cast<uintptr_t>(base) & 15) return set_error(-21, "TMA base address not 16-byte aligned");
  CUresult r = fn(map, dtype == B2U_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                  static_cast<cuuint32_t>(rank), const_cast<void*>(base), dims, strides, box, elem_strides,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(-22, "cuTensorMapEncodeTiled failed (CUresult %d; rank %d dims %llu,%llu box %u,%u)", (int)r, rank,
                     (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
  return 0;
}

}  // namespace b2u

namespace b2u {
static int g_options[8] = {0, 0, 0, 0, 0, 0, 0, 0};
int get_option(int key) { return (key >= 0 && key < 8) ? g_options[key] : 0; }
int current_device_index() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev < 0 || dev >= 64) ? 0 : dev;
}

int num_sms() {
  static int sms[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (sms[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    sms[dev] = v;
  }
  return sms[dev];
}
}  // namespace b2u

extern "C" int b2u_set_option(int32_t key, int32_t value) {
  if (key < 0 || key >= 8) return b2u::set_error(-1, "b2u_set_option: bad key %d", key);
  b2u::g_options[key] = value;
  return 0;
}

extern "C" int b2u_zero(void* ptr, int64_t bytes, b2u_stream_t stream) {
  cudaError_t e = cudaMemsetAsync(ptr, 0, static_cast<size_t>(bytes), static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return b2u::set_error(-11, "cudaMemsetAsync: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" const char* b2u_last_error(void) { return b2u::g_err; }
extern "C" int b2u_version(void) { return 1; }
extern "C" int64_t b2u_launch_count(void) { return b2u::g_launches.load(); }
