#include "common.cuh"

#include <stdarg.h>
#include <string.h>

#include <mutex>

namespace b200 {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cached = 0;
  if (cached) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  cached = n;
  return n;
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
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_nd_bf16(CUtensorMap* out, const void* base, int nd, const uint64_t* dims, const uint64_t* strides_elems,
                      const uint32_t* box, bool swizzle128) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled driver entry point unavailable (no CUDA driver / GPU)");
    return B200_ERR_NODEV;
  }
  B200_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer %p not 16-byte aligned", base);
  cuuint64_t gdim[5];
  cuuint64_t gstride[5];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < nd; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    B200_REQUIRE(box[i] >= 1 && box[i] <= 256, "TMA box dim %d = %u out of range", i, box[i]);
  }
  for (int i = 1; i < nd; ++i) {
    gstride[i - 1] = strides_elems[i] * 2;
    B200_REQUIRE((gstride[i - 1] & 15) == 0, "TMA stride %d (%llu bytes) not a multiple of 16", i,
                 (unsigned long long)gstride[i - 1]);
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)nd, const_cast<void*>(base), gdim, gstride, bdim,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult %d (nd=%d dims=%llu,%llu box=%u,%u)", (int)r, nd,
                   (unsigned long long)dims[0], (unsigned long long)(nd > 1 ? dims[1] : 0), box[0], nd > 1 ? box[1] : 0);
    return B200_ERR_CUDA;
  }
  return B200_OK;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
                      uint32_t box_rows) {
  uint64_t dims[2] = {cols, rows};
  uint64_t strides[2] = {1, ld};
  uint32_t box[2] = {box_cols, box_rows};
  return make_tmap_nd_bf16(out, base, 2, dims, strides, box, box_cols * 2 == 128);
}

}  // namespace b200

extern "C" const char* b200_last_error(void) { return b200::g_last_error; }

extern "C" int b200_abi_version(void) { return 1; }

// 0 when the current device is sm_100 (B200), B200_ERR_NODEV otherwise: product code calls this once and raises.
extern "C" int b200_device_check(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    b200::set_last_error("no CUDA device");
    return B200_ERR_NODEV;
  }
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    b200::set_last_error("device compute capability %d.%d is not sm_100 (B200)", major, minor);
    return B200_ERR_NODEV;
  }
  return B200_OK;
}
