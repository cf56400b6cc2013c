// Host-side helpers shared by the C-ABI translation units: error codes, TMA tensor-map encoding through the driver
// entry point (no link-time dependency on libcuda, so the library also loads on a GPU-less build box).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define B200_OK 0
#define B200_ERR_INVALID (-22)   // EINVAL: bad shape / alignment
#define B200_ERR_NODEV (-19)     // ENODEV: no sm_100 device / driver entry point missing
#define B200_ERR_CUDA (-5)       // EIO: CUDA runtime error (message kept in b200_last_error)

namespace b200 {

void set_last_error(const char* fmt, ...);

#define B200_CHECK_CUDA(expr)                                                                  \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      b200::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return B200_ERR_CUDA;                                                                    \
    }                                                                                          \
  } while (0)

#define B200_REQUIRE(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      b200::set_last_error(__VA_ARGS__); \
      return B200_ERR_INVALID;           \
    }                                    \
  } while (0)

int num_sms();

// bf16 2-D row-major tensor [rows, cols] with leading dimension ld (elements); box = {box_cols, box_rows}; 128B swizzle
// when box_cols * 2 == 128, else no swizzle.  Returns B200_OK or an error code.
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
                      uint32_t box_rows);
// generic N-d (<=4) bf16 map: dims[0] is the contiguous one; strides in elements for dims 1..n-1
int make_tmap_nd_bf16(CUtensorMap* out, const void* base, int nd, const uint64_t* dims, const uint64_t* strides_elems,
                      const uint32_t* box, bool swizzle128);

}  // namespace b200
