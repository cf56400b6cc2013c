// Decode-shaped nn.Linear: y[M, N] = x[M, K] W[N, K]^T with M <= 4 (one token per sequence in generate(),
// generation/utils.py decode loop -> LlamaAttention / LlamaMLP projections, models/llama/modeling_llama.py:174-176, :254-256,
// :280, :480).  At M <= 4 the layer is a stream over the weight matrix: N*K*2 bytes from HBM against 2*M*N*K FLOPs, i.e.
// <= 4 FLOP/byte -- HBM-bound by two orders of magnitude, so this is a CUDA-core kernel on purpose (a 128x256 tensor-core
// tile would occupy N/256 of the 148 SMs and leave the memory system idle; the tcgen05 GEMM stays the M > 4 path).
//
// One warp per output column n: the lane reads 16-byte pieces of W[n, :] (coalesced 512 B per warp access, each weight
// byte read exactly once), multiplies with the matching pieces of the M activation rows (M*K*2 bytes, L1/L2 resident,
// read through the read-only path) and the warp reduces with shuffles.  Algorithmic bytes: N*K*2 (+ M*K*2 + M*N*2).
#include <cuda_bf16.h>

#ifndef B200_HOST_EMU
#include "common.cuh"
#endif

namespace b200 {

constexpr int GEMV_WARPS = 8;

__device__ __forceinline__ void gemv_unpack8(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

template <int M>
__global__ void __launch_bounds__(GEMV_WARPS * 32)
gemv_bf16_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ W, __nv_bfloat16* __restrict__ y,
                 int N, int K, int ldx, int ldw, int ldy) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * GEMV_WARPS + warp;
  if (n >= N) return;
  const uint4* wrow = reinterpret_cast<const uint4*>(W + static_cast<size_t>(n) * ldw);
  float acc[M];
#pragma unroll
  for (int m = 0; m < M; ++m) acc[m] = 0.f;
  const int K8 = K / 8;
  // two independent 16-byte weight loads in flight per lane and iteration
  int c = lane;
  for (; c + 32 < K8; c += 64) {
    const uint4 w0 = wrow[c], w1 = wrow[c + 32];
    float fw0[8], fw1[8];
    gemv_unpack8(w0, fw0);
    gemv_unpack8(w1, fw1);
#pragma unroll
    for (int m = 0; m < M; ++m) {
      const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(m) * ldx);
      float fx0[8], fx1[8];
      gemv_unpack8(__ldg(xr + c), fx0);
      gemv_unpack8(__ldg(xr + c + 32), fx1);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[m] = fmaf(fw0[e], fx0[e], fmaf(fw1[e], fx1[e], acc[m]));
    }
  }
  for (; c < K8; c += 32) {
    float fw[8];
    gemv_unpack8(wrow[c], fw);
#pragma unroll
    for (int m = 0; m < M; ++m) {
      float fx[8];
      gemv_unpack8(__ldg(reinterpret_cast<const uint4*>(x + static_cast<size_t>(m) * ldx) + c), fx);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[m] = fmaf(fw[e], fx[e], acc[m]);
    }
  }
#pragma unroll
  for (int m = 0; m < M; ++m) {
    float a = acc[m];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) y[static_cast<size_t>(m) * ldy + n] = __float2bfloat16_rn(a);
  }
}

#ifndef B200_HOST_EMU
template <int M>
static int launch_gemv(const void* x, const void* W, void* y, int N, int K, int ldx, int ldw, int ldy, cudaStream_t stream) {
  gemv_bf16_kernel<M><<<(N + GEMV_WARPS - 1) / GEMV_WARPS, GEMV_WARPS * 32, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(W),
      reinterpret_cast<__nv_bfloat16*>(y), N, K, ldx, ldw, ldy);
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

#endif  // B200_HOST_EMU

}  // namespace b200

#ifndef B200_HOST_EMU
using namespace b200;

// y[M, N] = x[M, K] W[N, K]^T, 1 <= M <= 4, both operands K-major (x row stride ldx, W row stride ldw, elements).
extern "C" int b200_gemv_bf16(const void* x, const void* W, void* y, int M, int N, int K, int ldx, int ldw, int ldy,
                              cudaStream_t stream) {
  B200_REQUIRE(M >= 1 && M <= 4, "gemv: M=%d must be in 1..4", M);
  B200_REQUIRE(N > 0 && K > 0 && K % 8 == 0, "gemv: N=%d, K=%d (K must be a positive multiple of 8)", N, K);
  B200_REQUIRE(ldx % 8 == 0 && ldw % 8 == 0, "gemv: ldx / ldw must be multiples of 8 elements");
  B200_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
               "gemv: x and W must be 16B aligned");
  switch (M) {
    case 1: return launch_gemv<1>(x, W, y, N, K, ldx, ldw, ldy, stream);
    case 2: return launch_gemv<2>(x, W, y, N, K, ldx, ldw, ldy, stream);
    case 3: return launch_gemv<3>(x, W, y, N, K, ldx, ldw, ldy, stream);
    default: return launch_gemv<4>(x, W, y, N, K, ldx, ldw, ldy, stream);
  }
}
#endif  // B200_HOST_EMU
