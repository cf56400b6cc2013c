// RoPE cos / sin tables: LlamaRotaryEmbedding.forward (models/llama/modeling_llama.py:113-127; identical in Mistral,
// Gemma, Gemma2, Mixtral).  The reference runs six torch ops per forward (expand, fp32 outer product as a batched matmul
// with K = 1, transpose, cat, cos / sin, scale, cast); this is one launch with the same arithmetic in the same order:
//   angle = fp32(inv_freq[j]) * fp32(position)        (a K = 1 matmul is a single product: bit-identical)
//   cos   = bf16(cosf(angle) * attention_scaling)     (cosf / sinf: the functions torch's CUDA cos / sin call)
//   emb   = cat(freqs, freqs)  ->  column j and j + D/2 hold the same value
// so the tables are bit-exact against the reference's.  HBM-bound and tiny (2 * B * S * D * 2 B written per forward).
#ifndef B200_HOST_EMU
#include "common.cuh"
#endif

#include <cuda_bf16.h>
#include <math.h>

namespace b200 {

__global__ void rope_table_kernel(const float* __restrict__ inv_freq, const int64_t* __restrict__ pos,
                                  __nv_bfloat16* __restrict__ cos_t, __nv_bfloat16* __restrict__ sin_t, int rows, int half,
                                  float scaling) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * half) return;
  const int r = idx / half, j = idx - r * half;
  const float angle = inv_freq[j] * static_cast<float>(pos[r]);
  const __nv_bfloat16 c = __float2bfloat16_rn(cosf(angle) * scaling);
  const __nv_bfloat16 s = __float2bfloat16_rn(sinf(angle) * scaling);
  const size_t o = static_cast<size_t>(r) * (2 * half) + j;
  cos_t[o] = c;
  cos_t[o + half] = c;
  sin_t[o] = s;
  sin_t[o + half] = s;
}

}  // namespace b200

#ifndef B200_HOST_EMU
// inv_freq fp32 [D/2]; position_ids int64 [rows = B*S] (contiguous); cos / sin bf16 [rows, D].
extern "C" int b200_rope_table(const float* inv_freq, const int64_t* position_ids, void* cos_t, void* sin_t, int rows,
                               int D, float attention_scaling, cudaStream_t stream) {
  using namespace b200;
  B200_REQUIRE(D > 0 && D % 2 == 0, "rope_table: rotary dim %d must be even", D);
  B200_REQUIRE(rows >= 0, "rope_table: negative row count %d", rows);
  if (rows == 0) return B200_OK;
  const int half = D / 2;
  const long long total = static_cast<long long>(rows) * half;
  B200_REQUIRE(total < (1ll << 31), "rope_table: %lld elements exceed the 32-bit index range", total);
  const int threads = 256;
  const int grid = static_cast<int>((total + threads - 1) / threads);
  rope_table_kernel<<<grid, threads, 0, stream>>>(inv_freq, position_ids, reinterpret_cast<__nv_bfloat16*>(cos_t),
                                                  reinterpret_cast<__nv_bfloat16*>(sin_t), rows, half, attention_scaling);
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}
#endif  // B200_HOST_EMU
