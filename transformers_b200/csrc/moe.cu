// Mixture-of-experts token routing for the Mixtral experts path
// (MixtralExperts.forward models/mixtral/modeling_mixtral.py:69-93; grouped_mm_experts_forward integrations/moe.py:377-478):
// the reference sorts (token, k) pairs by expert (argsort + histc + cumsum), gathers the token rows, runs grouped GEMMs and
// un-permutes with index_add.  Here:
//   moe_count   : histogram of top_k_index over the experts (one atomic per entry, E <= 4096 bins in shared memory)
//   moe_scan    : exclusive prefix sum -> expert offsets (E is tiny)
//   moe_scatter : slot[t,k] = offset[e] + running cursor[e]  (order inside an expert is arbitrary; the final per-token
//                 sum over k is accumulated in fp32 in fixed k order -> deterministic)
//   moe_gather  : x_sorted[slot, :] = x[t, :]                          (16 B vector copies, warp per row)
//   moe_combine : out[t, :] = sum_k bf16(w[t,k] * y_sorted[slot[t,k], :])   (fp32 accumulate, bf16 out)
// The expert GEMMs themselves are b200_gemm_bf16 launches on contiguous row ranges of x_sorted.
#ifndef B200_HOST_EMU
#include "common.cuh"
#endif

#include <cuda_bf16.h>

namespace b200 {

__global__ void moe_count_kernel(const int64_t* __restrict__ idx, int* __restrict__ counts, int n, int E) {
#ifdef B200_HOST_EMU
  static int sh[4096];  // dynamic shared memory of the launch (E ints)
#else
  extern __shared__ int sh[];
#endif
  for (int e = threadIdx.x; e < E; e += blockDim.x) sh[e] = 0;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int64_t e = idx[i];
    if (e >= 0 && e < E) atomicAdd(&sh[e], 1);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += blockDim.x)
    if (sh[e]) atomicAdd(&counts[e], sh[e]);
}

__global__ void moe_scan_kernel(const int* __restrict__ counts, int* __restrict__ offsets, int* __restrict__ cursor, int E) {
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int e = 0; e < E; ++e) {
      offsets[e] = acc;
      cursor[e] = 0;
      acc += counts[e];
    }
    offsets[E] = acc;
  }
}

__global__ void moe_scatter_kernel(const int64_t* __restrict__ idx, const int* __restrict__ offsets, int* __restrict__ cursor,
                                   int* __restrict__ slot, int* __restrict__ token_of_slot, int n, int topk, int E) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t e = idx[i];
  if (e < 0 || e >= E) {
    slot[i] = -1;
    return;
  }
  const int s = offsets[e] + atomicAdd(&cursor[e], 1);
  slot[i] = s;
  token_of_slot[s] = i / topk;
}

__global__ void moe_gather_kernel(const uint4* __restrict__ x, const int* __restrict__ token_of_slot, uint4* __restrict__ xs,
                                  int nslots, int H8) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= nslots) return;
  const uint4* src = x + static_cast<size_t>(token_of_slot[row]) * H8;
  uint4* dst = xs + static_cast<size_t>(row) * H8;
  for (int c = lane; c < H8; c += 32) dst[c] = src[c];
}

__global__ void moe_combine_kernel(const uint4* __restrict__ ys, const int* __restrict__ slot, const float* __restrict__ w,
                                   uint4* __restrict__ out, int T, int topk, int H8) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (t >= T) return;
  for (int c = lane; c < H8; c += 32) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < topk; ++k) {
      const int s = slot[t * topk + k];
      if (s < 0) continue;
      const uint4 v = ys[static_cast<size_t>(s) * H8 + c];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
      // the reference multiplies in the activation dtype: bf16(weight) * bf16(out) -> bf16 (modeling_mixtral.py:90)
      const float wk = __bfloat162float(__float2bfloat16_rn(w[t * topk + k]));
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h[e]);
        acc[2 * e] += __bfloat162float(__float2bfloat16_rn(f.x * wk));
        acc[2 * e + 1] += __bfloat162float(__float2bfloat16_rn(f.y * wk));
      }
    }
    uint4 o;
    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) oh[e] = __floats2bfloat162_rn(acc[2 * e], acc[2 * e + 1]);
    out[static_cast<size_t>(t) * H8 + c] = o;
  }
}

}  // namespace b200

#ifndef B200_HOST_EMU
using namespace b200;

// top_k_index int64 [T, topk] -> counts[E] (zeroed by the caller), offsets[E+1], cursor[E] (scratch), slot[T*topk],
// token_of_slot[T*topk]
extern "C" int b200_moe_route(const int64_t* top_k_index, int* counts, int* offsets, int* cursor, int* slot,
                              int* token_of_slot, int T, int topk, int E, cudaStream_t stream) {
  B200_REQUIRE(E > 0 && E <= 4096 && topk > 0, "moe_route: bad E=%d topk=%d", E, topk);
  const int n = T * topk;
  if (n == 0) return B200_OK;
  int grid = (n + 255) / 256;
  if (grid > 1024) grid = 1024;
  moe_count_kernel<<<grid, 256, E * sizeof(int), stream>>>(top_k_index, counts, n, E);
  B200_CHECK_CUDA(cudaGetLastError());
  moe_scan_kernel<<<1, 32, 0, stream>>>(counts, offsets, cursor, E);
  B200_CHECK_CUDA(cudaGetLastError());
  moe_scatter_kernel<<<(n + 255) / 256, 256, 0, stream>>>(top_k_index, offsets, cursor, slot, token_of_slot, n, topk, E);
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_moe_gather(const void* x, const int* token_of_slot, void* x_sorted, int nslots, int H,
                               cudaStream_t stream) {
  B200_REQUIRE(H % 8 == 0, "moe_gather: H=%d must be a multiple of 8", H);
  if (nslots == 0) return B200_OK;
  moe_gather_kernel<<<(nslots * 32 + 255) / 256, 256, 0, stream>>>(reinterpret_cast<const uint4*>(x), token_of_slot,
                                                                   reinterpret_cast<uint4*>(x_sorted), nslots, H / 8);
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_moe_combine(const void* y_sorted, const int* slot, const float* weights, void* out, int T, int topk,
                                int H, cudaStream_t stream) {
  B200_REQUIRE(H % 8 == 0, "moe_combine: H=%d must be a multiple of 8", H);
  if (T == 0) return B200_OK;
  moe_combine_kernel<<<(T * 32 + 255) / 256, 256, 0, stream>>>(reinterpret_cast<const uint4*>(y_sorted), slot, weights,
                                                               reinterpret_cast<uint4*>(out), T, topk, H / 8);
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}
#endif  // B200_HOST_EMU
