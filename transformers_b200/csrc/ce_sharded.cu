// Vocabulary-sharded cross-entropy gradient (tensor-parallel lm_head without a logits all-gather): see
// functional.VocabParallelLossFn.  Reference semantics: ForCausalLMLoss (loss/loss_utils.py:32-70) on logits whose vocabulary
// dimension is split across ranks by the tp_plan entry "lm_head": "colwise..." (models/llama/modeling_llama.py:423).
// Kept in its own translation unit so that tests/emu can execute the kernel source on the host.
#include <cuda_bf16.h>

#ifndef B200_HOST_EMU
#include "common.cuh"
#endif

namespace b200 {

__device__ __forceinline__ void ces_unpack8(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 ces_pack8(const float (&f)[8]) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}

// Vocabulary-sharded variant (tensor-parallel lm_head without a logits gather): this rank holds V columns of every row.
// dlogits[row, v] = (exp(logit - lse_global[row]) - [v == target_local[row]]) * row_scale[row]
// target_local is the target's column inside this shard, or any value outside [0, V) when another rank owns it;
// row_scale is dloss / denom for valid rows and 0 for ignored ones.
__global__ void __launch_bounds__(1024)
ce_bwd_sharded_kernel(const __nv_bfloat16* __restrict__ logits, const int64_t* __restrict__ target_local,
                      const float* __restrict__ lse, const float* __restrict__ row_scale,
                      __nv_bfloat16* __restrict__ dlogits, int V, int ld, int ld_out) {
  const int row = blockIdx.x;
  const int64_t tgt = target_local[row];
  const float scale = row_scale[row];
  const float l = lse[row];
  const __nv_bfloat16* lrow = logits + static_cast<size_t>(row) * ld;
  __nv_bfloat16* drow = dlogits + static_cast<size_t>(row) * ld_out;
  const int V8 = V / 8;
  for (int c = threadIdx.x; c < V8; c += blockDim.x) {
    float f[8], o[8];
    ces_unpack8(__ldg(reinterpret_cast<const uint4*>(lrow) + c), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float p = __expf(f[e] - l);
      if (c * 8 + e == tgt) p -= 1.f;
      o[e] = p * scale;
    }
    *(reinterpret_cast<uint4*>(drow) + c) = ces_pack8(o);
  }
  for (int c = V8 * 8 + threadIdx.x; c < V; c += blockDim.x) {
    float p = __expf(__bfloat162float(lrow[c]) - l);
    if (c == tgt) p -= 1.f;
    drow[c] = __float2bfloat16_rn(p * scale);
  }
}

}  // namespace b200

#ifndef B200_HOST_EMU
using namespace b200;

#define B200_ALIGNED16(p) ((reinterpret_cast<uintptr_t>(p) & 15) == 0)

extern "C" int b200_ce_bwd_sharded(const void* logits, const int64_t* target_local, const float* lse_global,
                                   const float* row_scale, void* dlogits, int T, int V, int ld, int ld_out,
                                   cudaStream_t stream) {
  B200_REQUIRE(ld % 8 == 0 && ld_out % 8 == 0, "ce_bwd_sharded: rows must be 16B aligned");
  B200_REQUIRE(B200_ALIGNED16(logits) && B200_ALIGNED16(dlogits), "ce_bwd_sharded: pointers must be 16B aligned");
  if (T == 0 || V == 0) return B200_OK;
  ce_bwd_sharded_kernel<<<T, 1024, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(logits), target_local, lse_global,
                                                row_scale, reinterpret_cast<__nv_bfloat16*>(dlogits), V, ld, ld_out);
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}
#endif  // B200_HOST_EMU
