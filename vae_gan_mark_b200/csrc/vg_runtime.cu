// Host-side runtime pieces shared by all kernels: error string, device query, TMA descriptor encoding.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "vg_common.cuh"
#include "../../include/vaegan_b200.h"

namespace vg {

static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int g_sm_limit = 0;

int conv_sms() {
  const int n = num_sms();
  return (g_sm_limit > 0 && g_sm_limit < n) ? g_sm_limit : n;
}

int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev < 0 || dev >= 64) ? 0 : dev;
}

int num_sms() {
  static int cached[64] = {0};
  const int dev = current_device();
  if (cached[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cached[dev] = n > 0 ? n : 148;
  }
  return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                     const uint32_t* box) {
  EncodeTiledFn enc = get_encode();
  VG_CHECK(enc != nullptr, -3, "cuTensorMapEncodeTiled is not available from the driver");
  // cuTensorMapEncodeTiled is a driver entry point and needs a current context on the CALLING thread.  PyTorch's
  // autograd worker threads may not have bound one yet when the first thing a backward pass does is encode a map
  // (CUDA_ERROR_INVALID_CONTEXT); cudaFree(0) binds the device's primary context to this thread.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    VG_CUDA(cudaFree(nullptr));
    ctx_bound = true;
  }
  cuuint64_t gdim[5], gstride[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstride[i - 1] = strides_elems[i] * 2;  // bytes
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim,
                   gstride, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u]",
              static_cast<int>(r), rank, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
              (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0),
              (unsigned long long)(rank > 4 ? gdim[4] : 0), bdim[0], rank > 1 ? bdim[1] : 0, rank > 2 ? bdim[2] : 0,
              rank > 3 ? bdim[3] : 0, rank > 4 ? bdim[4] : 0);
    return -3;
  }
  return 0;
}

}  // namespace vg

extern "C" const char* vg_last_error(void) { return vg::g_err; }

extern "C" int vg_version(void) { return VG_API_VERSION; }

extern "C" int vg_set_conv_sm_limit(int n) {
  vg::g_sm_limit = n;
  return 0;
}

extern "C" unsigned long long vg_launch_count(void) { return vg::g_launches; }

extern "C" int vg_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  VG_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  VG_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return 0;
}
