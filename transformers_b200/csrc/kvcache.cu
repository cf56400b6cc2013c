// In-place KV-cache append for decode / generate().
//
// The reference's DynamicLayer.update (cache_utils.py:127-146) does torch.cat([cache, new], dim=-2): an O(context) copy
// of the whole K and V cache per layer per generated token (≈1.5 GB re-copied per decode step for Gemma-2-9B at 8.7k
// context).  Here the cache lives in a preallocated [B, Hkv, capacity, D] buffer and the new rows are
// written at their position: algorithmic traffic 2 * Hkv * D * 2 B per token per layer (4 KB for Llama-3-8B).
// One warp per (batch, head, new token) row; 16-byte vector copies; K and V in the same launch.
#ifndef B200_HOST_EMU
#include "common.cuh"
#endif

#include <cuda_bf16.h>

namespace b200 {

__global__ void kv_append_kernel(const __nv_bfloat16* __restrict__ k_new, const __nv_bfloat16* __restrict__ v_new,
                                 __nv_bfloat16* __restrict__ k_cache, __nv_bfloat16* __restrict__ v_cache, int B, int H,
                                 int q_len, int D8, int64_t ks_b, int64_t ks_h, int64_t ks_r, int64_t vs_b,
                                 int64_t vs_h, int64_t vs_r, int64_t cs_b, int64_t cs_h, int64_t cs_r, int offset) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int total = B * H * q_len;
  if (warp >= total) return;
  const int r = warp % q_len;
  const int h = (warp / q_len) % H;
  const int b = warp / (q_len * H);
  const uint4* ks = reinterpret_cast<const uint4*>(k_new + b * ks_b + h * ks_h + r * ks_r);
  const uint4* vs = reinterpret_cast<const uint4*>(v_new + b * vs_b + h * vs_h + r * vs_r);
  const int64_t dst = b * cs_b + h * cs_h + static_cast<int64_t>(offset + r) * cs_r;
  uint4* kd = reinterpret_cast<uint4*>(k_cache + dst);
  uint4* vd = reinterpret_cast<uint4*>(v_cache + dst);
  for (int c = lane; c < D8; c += 32) {
    kd[c] = ks[c];
    vd[c] = vs[c];
  }
}

}  // namespace b200

#ifndef B200_HOST_EMU
// k_new / v_new: [B, H, q_len, D] strided views (strides in elements: batch, head, row; unit inner stride);
// k_cache / v_cache: [B, H, capacity, D] with the given (batch, head, row) strides; rows [offset, offset + q_len) are
// written.  The caller guarantees offset + q_len <= capacity.
extern "C" int b200_kv_append(const void* k_new, const void* v_new, void* k_cache, void* v_cache, int B, int H,
                              int q_len, int D, int64_t ks_b, int64_t ks_h, int64_t ks_r, int64_t vs_b, int64_t vs_h,
                              int64_t vs_r, int64_t cs_b, int64_t cs_h, int64_t cs_r, int offset, int capacity,
                              cudaStream_t stream) {
  using namespace b200;
  B200_REQUIRE(D % 8 == 0, "kv_append: head_dim %d must be a multiple of 8", D);
  B200_REQUIRE(offset >= 0 && offset + q_len <= capacity, "kv_append: rows [%d, %d) exceed capacity %d", offset,
               offset + q_len, capacity);
  B200_REQUIRE(((ks_b | ks_h | ks_r | vs_b | vs_h | vs_r | cs_b | cs_h | cs_r) & 7) == 0,
               "kv_append: strides must be multiples of 8 elements");
  const int total = B * H * q_len;
  if (total == 0) return B200_OK;
  const int threads = 256;
  const int grid = (total * 32 + threads - 1) / threads;
  kv_append_kernel<<<grid, threads, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(k_new), reinterpret_cast<const __nv_bfloat16*>(v_new),
      reinterpret_cast<__nv_bfloat16*>(k_cache), reinterpret_cast<__nv_bfloat16*>(v_cache), B, H, q_len, D / 8, ks_b, ks_h,
      ks_r, vs_b, vs_h, vs_r, cs_b, cs_h, cs_r, offset);
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}
#endif  // B200_HOST_EMU
