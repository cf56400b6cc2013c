// Optimizer step right after the hot path: multi-tensor AdamW and global gradient-norm clipping.
//
// Reference call sites: Trainer._inner_training_loop clips (trainer.py:1783-1785 -> _clip_grad_norm :2538-2542 ->
// torch.nn.utils.clip_grad_norm_) and then calls optimizer.step() (:1788) on torch.optim.AdamW, which the reference selects
// for optim="adamw_torch"/"adamw_torch_fused" (trainer_optimizer.py:201-208).  torch is a third-party dependency of the
// reference (not vendored in /root/reference); the update below restates torch.optim.AdamW's published rule
// (decoupled weight decay, bias-corrected first/second moments, fp32 arithmetic on the loaded values) -- oracle/adamw_oracle.py
// is the CPU restatement, pinned against torch.optim.AdamW itself in tests/test_optim_cpu.py.
//
// All three kernels are HBM-bound streaming passes over every parameter tensor of a param group in ONE launch:
//   * tensor table (device, int64 [n_tensors][6]): {param*, grad*, exp_avg*, exp_avg_sq*, numel, fp32 master param* or 0}
//   * chunk map   (device, int32 [n_chunks][2]):   {tensor index, chunk index}; a chunk is OPT_CHUNK consecutive elements
// One CTA per chunk, 16-byte vector accesses whenever the four pointers are 16-byte aligned (always the case for torch
// allocations and for the row views of a packed weight buffer).  Algorithmic bytes per parameter: AdamW 14 B with bf16
// moments (read p,g,m,v + write p,m,v), 22 B with fp32 moments; norm 2 B; scale 4 B.
#include <cuda_bf16.h>

#ifndef B200_HOST_EMU  // tests/emu/cuda_emu.h provides the macros when the kernels are run on the host
#include "common.cuh"
#endif

namespace b200 {

constexpr int OPT_CHUNK = 32768;
constexpr int OPT_THREADS = 256;

struct AdamArgs {
  float lr, beta1, beta2, eps, weight_decay, step_size, inv_bc2_sqrt;
};

// approximate sqrt / division (<= 2 ulp in fp32): the stored results are rounded to bf16 (or feed a bf16 parameter), and
// IEEE sqrt + two IEEE divisions per element would make this 14 B/element stream ALU-bound (measured on the GLU kernels,
// profiles/README.md)
__device__ __forceinline__ float sqrt_approx(float x) {
#ifdef B200_HOST_EMU
  return sqrtf(x);
#else
  float y;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}

__device__ __forceinline__ void adamw_update(float& p, float g, float& m, float& v, const AdamArgs& a) {
  p -= a.lr * a.weight_decay * p;                         // decoupled weight decay
  m = m + (g - m) * (1.f - a.beta1);                      // exp_avg.lerp_(grad, 1 - beta1)
  v = a.beta2 * v + (1.f - a.beta2) * g * g;              // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrt_approx(v) * a.inv_bc2_sqrt + a.eps;  // (sqrt(v) / sqrt(1 - beta2^t)) + eps
  p -= a.step_size * __fdividef(m, denom);                // step_size = lr / (1 - beta1^t)
}

__device__ __forceinline__ void load8(const __nv_bfloat16* ptr, float (&f)[8]) {
  const uint4 v = *reinterpret_cast<const uint4*>(ptr);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void load8(const float* ptr, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(ptr);
  const float4 b = *reinterpret_cast<const float4*>(ptr + 4);
  f[0] = a.x, f[1] = a.y, f[2] = a.z, f[3] = a.w, f[4] = b.x, f[5] = b.y, f[6] = b.z, f[7] = b.w;
}
__device__ __forceinline__ void store8(__nv_bfloat16* ptr, const float (&f)[8]) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<uint4*>(ptr) = v;
}
__device__ __forceinline__ void store8(float* ptr, const float (&f)[8]) {
  *reinterpret_cast<float4*>(ptr) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(ptr + 4) = make_float4(f[4], f[5], f[6], f[7]);
}
__device__ __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ void from_f(__nv_bfloat16& d, float x) { d = __float2bfloat16_rn(x); }
__device__ __forceinline__ void from_f(float& d, float x) { d = x; }

__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// MASTER: the update runs on an fp32 master copy of the parameter (table column 5) and the bf16 parameter is its rounding --
// bf16 alone drops updates smaller than half an ulp (|dp| < 2^-9 |p|), which is most of them late in training.
template <typename ST, bool MASTER>
__global__ void __launch_bounds__(OPT_THREADS)
adamw_multi_kernel(const int64_t* __restrict__ table, const int2* __restrict__ chunks, AdamArgs a,
                   const float* __restrict__ grad_scale) {
  const int2 ch = chunks[blockIdx.x];
  const int64_t* row = table + static_cast<int64_t>(ch.x) * 6;
  __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(row[0]);
  const __nv_bfloat16* g = reinterpret_cast<const __nv_bfloat16*>(row[1]);
  ST* m = reinterpret_cast<ST*>(row[2]);
  ST* v = reinterpret_cast<ST*>(row[3]);
  const int64_t n = row[4];
  float* master = MASTER ? reinterpret_cast<float*>(row[5]) : nullptr;
  const int64_t start = static_cast<int64_t>(ch.y) * OPT_CHUNK;
  const int64_t end = (start + OPT_CHUNK < n) ? start + OPT_CHUNK : n;
  const float gs = grad_scale ? *grad_scale : 1.f;
  int64_t done = start;
  if (aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v) && (!MASTER || aligned16(master))) {  // OPT_CHUNK % 8 == 0
    const int64_t vec_end = start + ((end - start) / 8) * 8;
    for (int64_t i = start + threadIdx.x * 8; i < vec_end; i += OPT_THREADS * 8) {
      float fp[8], fg[8], fm[8], fv[8];
      if (MASTER) load8(master + i, fp);
      else load8(p + i, fp);
      load8(g + i, fg);
      load8(m + i, fm);
      load8(v + i, fv);
#pragma unroll
      for (int e = 0; e < 8; ++e) adamw_update(fp[e], fg[e] * gs, fm[e], fv[e], a);
      if (MASTER) store8(master + i, fp);
      store8(p + i, fp);
      store8(m + i, fm);
      store8(v + i, fv);
    }
    done = vec_end;
  }
  for (int64_t i = done + threadIdx.x; i < end; i += OPT_THREADS) {
    float fp = MASTER ? master[i] : __bfloat162float(p[i]), fm = to_f(m[i]), fv = to_f(v[i]);
    adamw_update(fp, __bfloat162float(g[i]) * gs, fm, fv, a);
    if (MASTER) master[i] = fp;
    p[i] = __float2bfloat16_rn(fp);
    from_f(m[i], fm);
    from_f(v[i], fv);
  }
}

__device__ __forceinline__ float block_sum(float x, float* smem) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) smem[wid] = x;
  __syncthreads();
  x = (threadIdx.x < (blockDim.x >> 5)) ? smem[threadIdx.x] : 0.f;
  if (wid == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  }
  return x;  // valid in thread 0
}

// partial[chunk] = sum of grad^2 over the chunk (fp32); deterministic (no atomics)
__global__ void __launch_bounds__(OPT_THREADS)
grad_sq_norm_kernel(const int64_t* __restrict__ table, const int2* __restrict__ chunks, float* __restrict__ partial) {
  __shared__ float smem[32];
  const int2 ch = chunks[blockIdx.x];
  const int64_t* row = table + static_cast<int64_t>(ch.x) * 6;
  const __nv_bfloat16* g = reinterpret_cast<const __nv_bfloat16*>(row[1]);
  const int64_t n = row[4];
  const int64_t start = static_cast<int64_t>(ch.y) * OPT_CHUNK;
  const int64_t end = (start + OPT_CHUNK < n) ? start + OPT_CHUNK : n;
  float acc = 0.f;
  int64_t done = start;
  if (aligned16(g)) {
    const int64_t vec_end = start + ((end - start) / 8) * 8;
    for (int64_t i = start + threadIdx.x * 8; i < vec_end; i += OPT_THREADS * 8) {
      float f[8];
      load8(g + i, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc += f[e] * f[e];
    }
    done = vec_end;
  }
  for (int64_t i = done + threadIdx.x; i < end; i += OPT_THREADS) {
    const float f = __bfloat162float(g[i]);
    acc += f * f;
  }
  acc = block_sum(acc, smem);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// out[0] = total L2 norm, out[1] = clip coefficient min(1, max_norm / (norm + 1e-6)) (1 when max_norm <= 0)
// (torch.nn.utils.clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to 1.0)
__global__ void __launch_bounds__(1024) grad_norm_finish_kernel(const float* __restrict__ partial, int n, float max_norm,
                                                                float* __restrict__ out) {
  __shared__ double sm[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += static_cast<double>(partial[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = sm[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) {
      const float norm = static_cast<float>(sqrt(acc));
      out[0] = norm;
      const float coef = max_norm > 0.f ? max_norm / (norm + 1e-6f) : 1.f;
      out[1] = coef < 1.f ? coef : 1.f;
    }
  }
}

// grad *= *coef in place (skipped entirely when the coefficient is 1: the common no-clip case costs no HBM traffic)
__global__ void __launch_bounds__(OPT_THREADS)
grad_scale_kernel(const int64_t* __restrict__ table, const int2* __restrict__ chunks, const float* __restrict__ coef) {
  const float c = *coef;
  if (c == 1.f) return;
  const int2 ch = chunks[blockIdx.x];
  const int64_t* row = table + static_cast<int64_t>(ch.x) * 6;
  __nv_bfloat16* g = reinterpret_cast<__nv_bfloat16*>(row[1]);
  const int64_t n = row[4];
  const int64_t start = static_cast<int64_t>(ch.y) * OPT_CHUNK;
  const int64_t end = (start + OPT_CHUNK < n) ? start + OPT_CHUNK : n;
  int64_t done = start;
  if (aligned16(g)) {
    const int64_t vec_end = start + ((end - start) / 8) * 8;
    for (int64_t i = start + threadIdx.x * 8; i < vec_end; i += OPT_THREADS * 8) {
      float f[8];
      load8(g + i, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] *= c;
      store8(g + i, f);
    }
    done = vec_end;
  }
  for (int64_t i = done + threadIdx.x; i < end; i += OPT_THREADS) g[i] = __float2bfloat16_rn(__bfloat162float(g[i]) * c);
}

}  // namespace b200

#ifndef B200_HOST_EMU
using namespace b200;

extern "C" int b200_optim_chunk_elems(void) { return OPT_CHUNK; }

// state_is_fp32: bit 0 = the moments are fp32 (else bf16), bit 1 = table column 5 holds fp32 master parameters
extern "C" int b200_adamw_step(const int64_t* tensor_table, const int32_t* chunk_map, int n_chunks, int state_is_fp32,
                               float lr, float beta1, float beta2, float eps, float weight_decay, float bias_correction1,
                               float bias_correction2_sqrt, const float* grad_scale, cudaStream_t stream) {
  B200_REQUIRE(n_chunks >= 0, "adamw_step: negative chunk count");
  B200_REQUIRE(bias_correction1 > 0.f && bias_correction2_sqrt > 0.f, "adamw_step: bias corrections must be positive (step >= 1)");
  B200_REQUIRE(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f && lr >= 0.f && weight_decay >= 0.f,
               "adamw_step: hyper-parameters out of range");
  if (n_chunks == 0) return B200_OK;
  AdamArgs a{lr, beta1, beta2, eps, weight_decay, lr / bias_correction1, 1.f / bias_correction2_sqrt};
  const int2* chunks = reinterpret_cast<const int2*>(chunk_map);
  switch (state_is_fp32 & 3) {
    case 0: adamw_multi_kernel<__nv_bfloat16, false><<<n_chunks, OPT_THREADS, 0, stream>>>(tensor_table, chunks, a, grad_scale); break;
    case 1: adamw_multi_kernel<float, false><<<n_chunks, OPT_THREADS, 0, stream>>>(tensor_table, chunks, a, grad_scale); break;
    case 2: adamw_multi_kernel<__nv_bfloat16, true><<<n_chunks, OPT_THREADS, 0, stream>>>(tensor_table, chunks, a, grad_scale); break;
    default: adamw_multi_kernel<float, true><<<n_chunks, OPT_THREADS, 0, stream>>>(tensor_table, chunks, a, grad_scale); break;
  }
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_grad_norm(const int64_t* tensor_table, const int32_t* chunk_map, int n_chunks, float* partial_ws,
                              float max_norm, float* out2, cudaStream_t stream) {
  B200_REQUIRE(n_chunks >= 0, "grad_norm: negative chunk count");
  if (n_chunks > 0) {
    grad_sq_norm_kernel<<<n_chunks, OPT_THREADS, 0, stream>>>(tensor_table, reinterpret_cast<const int2*>(chunk_map), partial_ws);
    B200_CHECK_CUDA(cudaGetLastError());
  }
  grad_norm_finish_kernel<<<1, 1024, 0, stream>>>(partial_ws, n_chunks, max_norm, out2);
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_grad_scale(const int64_t* tensor_table, const int32_t* chunk_map, int n_chunks, const float* coef,
                               cudaStream_t stream) {
  B200_REQUIRE(n_chunks >= 0, "grad_scale: negative chunk count");
  if (n_chunks == 0) return B200_OK;
  grad_scale_kernel<<<n_chunks, OPT_THREADS, 0, stream>>>(tensor_table, reinterpret_cast<const int2*>(chunk_map), coef);
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}
#endif  // B200_HOST_EMU
