// b200seg — host utilities: thread-local error string, arch check, TMA descriptor encoding.
#include <stdarg.h>
#include <stdlib.h>
#include <map>
#include <mutex>
#include <string>

#include "common.cuh"

namespace b2 {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return B2_ERR_CUDA;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ---------------------------------------------------------------------------------------------
// environment switches, read once
// ---------------------------------------------------------------------------------------------
static std::mutex g_env_mu;
static std::map<std::string, int> g_env_cache;

int env_switch(const char* name, int dflt) {
  std::lock_guard<std::mutex> lk(g_env_mu);
  auto it = g_env_cache.find(name);
  if (it != g_env_cache.end()) return it->second == INT32_MIN ? dflt : it->second;
  const char* v = getenv(name);
  const int val = v ? atoi(v) : INT32_MIN;       // INT32_MIN = "not set": the caller's default applies
  g_env_cache.emplace(name, val);
  return v ? val : dflt;
}

// ---------------------------------------------------------------------------------------------
// deterministic-reduction workspace
// ---------------------------------------------------------------------------------------------
static std::mutex g_det_mu;
static double* g_det_ws[64] = {nullptr};       // per device
static long long g_det_doubles[64] = {0};

static bool det_lookup(double** ws, long long* n) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
  std::lock_guard<std::mutex> lk(g_det_mu);
  *ws = g_det_ws[dev];
  *n = g_det_doubles[dev];
  return *ws != nullptr;
}

bool det_enabled() {
  double* ws;
  long long n;
  return det_lookup(&ws, &n);
}

int det_begin(DetBuf* out, long long rows, int n, cudaStream_t stream, long long keep) {
  out->partial = nullptr;
  out->n = n;
  double* ws;
  long long cap;
  if (!det_lookup(&ws, &cap)) return B2_OK;
  const long long need = rows * (long long)n;
  B2_REQUIRE(keep + need <= cap, B2_ERR_WORKSPACE,
             "deterministic workspace too small: need %lld B, have %lld B (b2_set_deterministic)",
             (keep + need) * 8, cap * 8);
  out->partial = ws + keep;
  B2_CHECK_CUDA(cudaMemsetAsync(out->partial, 0, (size_t)need * sizeof(double), stream));
  return B2_OK;
}

template <typename T>
__global__ void det_finish_kernel(const double* __restrict__ partial, long long rows, int row_stride, int count,
                                  T* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  double s = 0.0;
  for (long long r = 0; r < rows; ++r) s += partial[r * row_stride + i];     // fixed order: bit-reproducible
  dst[i] = (T)((double)dst[i] + s);
}

int det_finish(const double* partial, long long rows, int row_stride, int count, double* dst, cudaStream_t stream) {
  if (count <= 0) return B2_OK;
  det_finish_kernel<double><<<(count + 127) / 128, 128, 0, stream>>>(partial, rows, row_stride, count, dst);
  B2_LAUNCH_CHECK();
  return B2_OK;
}
int det_finish(const double* partial, long long rows, int row_stride, int count, float* dst, cudaStream_t stream) {
  if (count <= 0) return B2_OK;
  det_finish_kernel<float><<<(count + 127) / 128, 128, 0, stream>>>(partial, rows, row_stride, count, dst);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz,
                     const uint32_t* elem_strides) {
  EncodeTiledFn fn = get_encode_fn();
  B2_REQUIRE(fn != nullptr, B2_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no driver?)");
  B2_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, B2_ERR_ALIGN, "TMA base pointer not 16B aligned");
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = elem_strides ? elem_strides[i] : 1;
    if (i > 0) {
      gstr[i - 1] = strides_bytes[i];
      B2_REQUIRE((strides_bytes[i] & 15) == 0, B2_ERR_ALIGN, "TMA stride %d = %llu B not a multiple of 16", i,
                 (unsigned long long)strides_bytes[i]);
    }
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr,
                  bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u]", (int)r,
              rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
              rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return B2_ERR_CUDA;
  }
  return B2_OK;
}

}  // namespace b2

extern "C" {

const char* b2_last_error(void) { return b2::g_err; }
int b2_abi_version(void) { return B2_ABI_VERSION; }
int b2_num_sms(void) { return b2::num_sms(); }

int b2_set_deterministic(void* workspace, int64_t bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return b2::cuda_fail(e, "cudaGetDevice");
  B2_REQUIRE(dev >= 0 && dev < 64, B2_ERR_SHAPE, "device index %d out of range", dev);
  B2_REQUIRE(workspace == nullptr || (bytes >= (1 << 20) && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0),
             B2_ERR_WORKSPACE, "deterministic workspace must be 16 B aligned and >= 1 MiB");
  std::lock_guard<std::mutex> lk(b2::g_det_mu);
  b2::g_det_ws[dev] = static_cast<double*>(workspace);
  b2::g_det_doubles[dev] = workspace ? bytes / 8 : 0;
  return B2_OK;
}

int b2_get_deterministic(void) { return b2::det_enabled() ? 1 : 0; }

int b2_reload_env(void) {
  std::lock_guard<std::mutex> lk(b2::g_env_mu);
  b2::g_env_cache.clear();
  return B2_OK;
}

int b2_arch_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return b2::cuda_fail(e, "cudaGetDevice");
  int major = 0, minor = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return b2::cuda_fail(e, "cudaDeviceGetAttribute");
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10 || minor != 0) {
    b2::set_error("libb200seg is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
    return B2_ERR_ARCH;
  }
  return B2_OK;
}

}  // extern "C"
