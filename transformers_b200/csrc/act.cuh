// Gated-MLP activation shared by the stand-alone GLU kernels (elementwise.cu) and the GLU epilogue of the CTA-pair GEMM
// (gemm2.cu), so that the fused and the unfused path produce bit-identical results.
// reference: LlamaMLP.forward models/llama/modeling_llama.py:174-176: h = bf16( bf16(act(g)) * u ), g and u being the bf16
// outputs of the gate / up projections; act = silu (activations.py:92-103) or gelu(approximate="tanh") (:30-49, Gemma).
// One transcendental per element: sigmoid(g) = rcp(1 + 2^(-g*log2e)) (MUFU.EX2 + MUFU.RCP) or tanh.approx (MUFU.TANH).
#pragma once
#include <cuda_bf16.h>

namespace b200 {

__device__ __forceinline__ float act_bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ float fast_tanhf(float x) {
#ifdef B200_HOST_EMU
  return tanhf(x);
#else
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}
__device__ __forceinline__ void act_fwd_grad(float g, int gelu, float& act, float& dact) {
  if (!gelu) {
    const float s = __frcp_rn(1.0f + __expf(-g));
    act = g * s;
    dact = s * (1.0f + g * (1.0f - s));
  } else {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    const float g2 = g * g;
    const float t = fast_tanhf(k0 * g * (1.0f + k1 * g2));
    act = 0.5f * g * (1.0f + t);
    dact = 0.5f * (1.0f + t) + 0.5f * g * (1.0f - t * t) * k0 * (1.0f + 3.0f * k1 * g2);
  }
}
__device__ __forceinline__ float act_fwd(float g, int gelu) {
  if (!gelu) return g * __frcp_rn(1.0f + __expf(-g));
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  return 0.5f * g * (1.0f + fast_tanhf(k0 * g * (1.0f + k1 * g * g)));
}
// h = act(g) * u with the reference's rounding points; g, u already bf16-representable; the caller rounds h to bf16
__device__ __forceinline__ float glu_value(float g, float u, int gelu) { return act_bf16_round(act_fwd(g, gelu)) * u; }

}  // namespace b200
