// Shared pieces of the tcgen05 attention kernels (forward, dK/dV backward, dQ backward).
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace b200 {

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));  // not volatile: pure, free to schedule
  return y;
}
__device__ __forceinline__ float fast_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 1-D bulk copy global -> shared with mbarrier completion (16-byte aligned, size multiple of 16)
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Mask description shared by all three kernels.  q row i sits at kv position i + (Skv - Sq) (bottom-right aligned).
struct AttnMask {
  int Sq, Skv;
  int causal;
  int window;           // 0 = off; otherwise kv_idx > q_pos - window
  const int* kv_start;  // optional [B]
  const int* kv_end;    // optional [B]
};

// 4-D tensor map over strided [B, S, h, D] storage: dims {D, S, h, B}; box {64, box_rows, 1, 1}, 128B swizzle
static inline int make_qkv_tmap(CUtensorMap* tm, const void* ptr, int D, int S, int H, int B, int64_t batch_stride,
                                int64_t row_stride, int64_t head_stride, int box_rows) {
  uint64_t dims[4] = {(uint64_t)D, (uint64_t)S, (uint64_t)H, (uint64_t)B};
  uint64_t strides[4] = {1, (uint64_t)row_stride, (uint64_t)head_stride, (uint64_t)batch_stride};
  uint32_t box[4] = {64, (uint32_t)box_rows, 1, 1};
  return make_tmap_nd_bf16(tm, ptr, 4, dims, strides, box, true);
}

}  // namespace b200
